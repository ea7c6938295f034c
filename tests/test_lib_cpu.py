"""CPU: the C-ABI library builds, loads and exports every symbol include/cris_b200.h declares (no compute calls),
and the ctypes struct mirror matches the C struct layout."""
import ctypes as C
import re
from pathlib import Path

from cris.pytorch_b200 import _lib, build

REPO = Path(__file__).resolve().parent.parent


def test_library_builds_and_exports_header_symbols():
    lib_path = build.build_library()
    assert lib_path.exists()
    L = _lib.lib()
    header = (REPO / "include" / "cris_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cris_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in cris_b200.h but not exported"
    assert set(_lib.exported_symbols()) >= declared - {"cris_gemm_args"}
    assert L.cris_abi_version() == 1
    assert L.cris_gemm_args_size() == C.sizeof(_lib.GemmArgs)


def test_no_gpu_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.device_check()
