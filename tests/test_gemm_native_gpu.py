"""GPU: the native differential test of the tcgen05 GEMM core (tests/native/gemm_selftest.cu): every case runs the
persistent tensor-core kernel and the SIMT restatement (gemm_ref.cu) through the same C-ABI entry point on the same
operands — all operand majorness combinations, 3x3 tap modes (forward / dgrad / wgrad), two-level batches, split-K
(explicit and planned), bias / ReLU / QuickGELU / residual / border-mask / column-statistics epilogues, ragged
M/N/K — and compares D (and the statistics) element-wise; a sample is also checked against a host loop.
Each case runs in its own process so that a trapped kernel cannot poison the others."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
BIN = Path(__file__).resolve().parent / "native" / "gemm_selftest"


def test_native_gemm_selftest_all_cases():
    if not BIN.exists():
        pytest.skip("tests/native/gemm_selftest is not built (python -m cris.pytorch_b200.build --selftest)")
    n = int(subprocess.run([str(BIN), "list"], capture_output=True, text=True, timeout=120).stdout.strip())
    assert n >= 40
    failed = []
    for i in range(n):
        r = subprocess.run([str(BIN), str(i)], capture_output=True, text=True, timeout=180)
        if r.returncode != 0 or "FAIL" in r.stdout:
            failed.append((i, r.returncode, r.stdout[-400:]))
    assert not failed, failed
