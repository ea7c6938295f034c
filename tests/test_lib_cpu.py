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
    assert L.cris_abi_version() == 2
    assert L.cris_gemm_args_size() == C.sizeof(_lib.GemmArgs)


def test_no_gpu_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.device_check()


def _plan(M, N, K, taps=1, splits=0, accumulate=1, a_mn=1, b_mn=1):
    L = _lib.lib()
    g = _lib.GemmArgs()
    g.M, g.N, g.K, g.batch, g.batch_inner = M, N, K, 1, 1
    g.a_mn, g.b_mn, g.d_fp32, g.accumulate, g.splits = a_mn, b_mn, int(bool(accumulate)), accumulate, splits
    g.tap_mode, g.taps = (2, taps) if taps > 1 else (0, 1)
    bn, sp = C.c_int(), C.c_int()
    assert L.cris_gemm_plan(C.byref(g), C.byref(bn), C.byref(sp)) == 0, L.cris_last_error()
    return bn.value, sp.value


def test_split_k_plan_fills_whole_waves():
    """Host logic of the wgrad split-K planner (no GPU: 148 SMs assumed): the work units of every fp32-accumulating
    GEMM of the cris_r50 backward fill at least 85 % of the waves they occupy (>= 97 % for all but one shape), never split below 4 K-blocks, and an
    explicit split count is respected."""
    sms = 148
    shapes = [(256, 512, 719104, 9), (512, 512, 186624, 9), (512, 512, 50176, 9), (64, 64, 719104, 9),
              (128, 128, 719104, 9), (256, 256, 186624, 9), (512, 1024, 50176, 9), (32, 32, 2822400, 9),
              (2048, 512, 43264, 1), (512, 2048, 43264, 1), (512, 512, 43264, 1), (1024, 2048, 12544, 1),
              (64, 256, 719104, 1)]
    for M, N, K, taps in shapes:
        bn, sp = _plan(M, N, K, taps)
        assert bn in (64, 128, 256) and sp >= 1
        tiles = ((M + 127) // 128) * ((N + bn - 1) // bn) * taps
        units = tiles * sp
        waves = (units + sms - 1) // sms
        assert units / (waves * sms) >= 0.85, (M, N, K, taps, bn, sp, units)
        assert sp == 1 or (K + 63) // 64 // sp >= 4, (M, N, K, sp)
        assert bn == 64 or N > bn // 2
    assert _plan(512, 512, 43264, 1, splits=5)[1] == 5
    assert _plan(43264, 512, 512, 1, splits=1, accumulate=0, a_mn=0, b_mn=0) == (256, 1)
