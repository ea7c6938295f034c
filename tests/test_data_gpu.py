"""GPU: DevicePrefetcher (host->device staging one batch ahead, the wrapper around the reference's loader loop
engine/engine.py:40-46) yields every batch once, in order, bit-identical, on the device."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_yields_all_batches_in_order():
    from cris.pytorch_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(4, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 100, (4, 17), generator=g),
                {"mask": torch.rand(4, 32, 32, generator=g), "note": "kept"}) for _ in range(5)]
    pf = DevicePrefetcher(batches)
    assert len(pf) == 5
    seen = 0
    for (img, word, extra), (ri, rw, re) in zip(pf, batches):
        assert img.is_cuda and word.is_cuda and extra["mask"].is_cuda and extra["note"] == "kept"
        y = img * 2.0  # consume on the compute stream
        assert torch.equal(img.cpu(), ri) and torch.equal(word.cpu(), rw) and torch.equal(extra["mask"].cpu(), re["mask"])
        assert torch.equal(y.cpu(), ri * 2.0)
        assert img.cuda(non_blocking=True) is img  # the reference loop's .cuda() calls become no-ops
        seen += 1
    assert seen == 5
    assert list(DevicePrefetcher([])) == []
