"""GPU: the NVLink peer-memory statistics exchange (csrc/peer.cu) — the collective inside SyncBatchNorm
(reference train.py:97-98) done by ONE kernel per site.  Two processes map each other's IPC buffer; on a box with
two GPUs they sit on different devices, on a single-GPU box both use cuda:0 (the driver time-slices the two
spinning kernels), which exercises the same handshake.  Results must equal the rank-ordered fp32 sum exactly,
eagerly and when replayed from a CUDA graph."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _vec(rank, it, slot, n):
    g = torch.Generator().manual_seed(1000 * it + 10 * slot + rank)
    return torch.randn(n, generator=g, dtype=torch.float32)


def _expected(world, it, slot, n):
    s = torch.zeros(n, dtype=torch.float32)
    for r in range(world):  # rank order, fp32: the kernel's summation order
        s = s + _vec(r, it, slot, n)
    return s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CRIS_B200_PEER_TIMEOUT_S="30")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cris.pytorch_b200.peer import PeerExchange, MAX_SLOTS, SLOT_FLOATS
        dev = torch.device("cuda", rank % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        ex = PeerExchange(dev)
        if not ex.ok:
            q.put((rank, "unavailable", ex.error))
            return
        bad = 0
        sites = [(0, 1), (1, 130), (7, 2048), (MAX_SLOTS - 1, SLOT_FLOATS)]
        for it in range(5):  # odd and even epochs: both parities of every slot
            for slot, n in sites:
                t = _vec(rank, it, slot, n).to(dev)
                ex.allreduce(slot, t)
                bad += int(not torch.equal(t.cpu(), _expected(world, it, slot, n)))
        # replay from a CUDA graph: three sites per replay, inputs refreshed between replays
        bufs = [torch.zeros(n, device=dev) for _, n in sites[:3]]
        outs = [torch.zeros(n, device=dev) for _, n in sites[:3]]
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for (slot, n), b, o in zip(sites[:3], bufs, outs):
                    o.copy_(b)
                    ex.allreduce(slot, o)
        for it in range(5, 9):
            for (slot, n), b in zip(sites[:3], bufs):
                b.copy_(_vec(rank, it, slot, n))
            g.replay()
            torch.cuda.synchronize(dev)
            for (slot, n), o in zip(sites[:3], outs):
                bad += int(not torch.equal(o.cpu(), _expected(world, it, slot, n)))
        # fused SyncBatchNorm sites (peer_bn_sync_kernel): partials -> push exchange -> coefficients, eagerly and from a
        # graph, more sites than the ring has slots; reference = the same arithmetic in torch on the rank-ordered sums
        for it in range(40):
            C = (8, 64, 256, 2048)[it % 4]
            nt = (1, 5, 64)[it % 3]
            parts = [torch.randn(nt, 2, C, generator=torch.Generator().manual_seed(7000 + 10 * it + r)) for r in range(world)]
            for p_ in parts:
                p_[:, 1].abs_()
                p_[:, 1] += 50.0 * nt  # keep the variance positive
            gam, bet = torch.rand(C) + 0.5, torch.randn(C)
            rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
            coef = torch.zeros(4 * C, device=dev)
            count = 100.0 * world
            ex.bn_sync_fwd(it, parts[rank].to(dev).contiguous(), nt, C, count, gam.to(dev), bet.to(dev), 1e-5, 0.1, rm, rv, coef)
            tot = torch.zeros(2, C)
            for r in range(world):
                tot = tot + parts[r].sum(0)
            mean = (tot[0].double() / count)
            var = (tot[1].double() / count - mean * mean).clamp_min(0)
            inv = torch.rsqrt(var.float() + 1e-5)
            exp = torch.cat([gam * inv, bet - mean.float() * gam * inv, mean.float(), inv])
            got = coef.cpu()
            bad += int(not torch.allclose(got, exp, rtol=2e-4, atol=1e-5))
            bad += int(not torch.allclose(rm.cpu(), 0.1 * mean.float(), rtol=2e-4, atol=1e-6))
            # backward flavour: local sums -> parameter gradients, global sums out
            bs, g0, g1 = torch.zeros(2 * C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
            ex.bn_sync_bwd(1000 + it, parts[rank].to(dev).contiguous(), nt, C, bs, g0, g1)
            loc = parts[rank].sum(0)
            bad += int(not torch.allclose(g0.cpu(), loc[0], rtol=1e-4, atol=1e-4))
            bad += int(not torch.allclose(g1.cpu(), loc[1], rtol=1e-4, atol=1e-4))
            bad += int(not torch.allclose(bs.cpu(), tot.reshape(-1), rtol=1e-4, atol=1e-4))
        dist.barrier()
        ex.close()
        q.put((rank, "ok", bad))
    except Exception as e:  # surface worker failures instead of a queue timeout
        q.put((rank, "error", repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_peer_allreduce_two_processes():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    if any(r[1] == "unavailable" for r in res):
        pytest.skip(f"CUDA IPC peer mapping unavailable on this box: {res}")
    assert all(r[1] == "ok" and r[2] == 0 for r in res), res
    assert all(p.exitcode == 0 for p in procs)
