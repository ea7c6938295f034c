"""In-tree build of libcris_b200.so (hand-written sm_100a kernels + C ABI) with nvcc.

The built library lives next to the sources (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  `python -m cris.pytorch_b200.build` or `__graft_entry__.build()` runs it.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
REPO = HERE.parent.parent
LIB_PATH = HERE / "libcris_b200.so"
OBJ_DIR = HERE / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libcris_b200.so cannot be built")
    return nvcc


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: Path, obj: Path, log: Path) -> None:
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(REPO / "include"), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")


def build_library(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [REPO / "include" / "cris_b200.h"]
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_digest = _digest(headers)
    jobs = []
    for src in sources:
        obj = OBJ_DIR / (src.stem + ".o")
        stamp = OBJ_DIR / (src.stem + ".stamp")
        want = _digest([src]) + hdr_digest
        if force or not obj.exists() or not stamp.exists() or stamp.read_text() != want:
            jobs.append((src, obj, stamp, want))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = {ex.submit(_compile_one, s, o, OBJ_DIR / (s.stem + ".log")): (s, o, st, w) for s, o, st, w in jobs}
            for f in concurrent.futures.as_completed(futs):
                s, o, st, w = futs[f]
                f.result()
                st.write_text(w)
                if verbose:
                    print(f"[build] compiled {s.name}", file=sys.stderr)
    objs = [OBJ_DIR / (s.stem + ".o") for s in sources]
    if jobs or not LIB_PATH.exists():
        cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"[build] linked {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


def build_selftest() -> Path:
    """tests/native/gemm_selftest: the pure-CUDA differential test of the GEMM core."""
    lib = build_library()
    src = REPO / "tests" / "native" / "gemm_selftest.cu"
    ref = REPO / "tests" / "native" / "gemm_ref.cu"  # SIMT restatement: test-only, never part of the product library
    out = REPO / "tests" / "native" / "gemm_selftest"
    if out.exists() and out.stat().st_mtime > max(src.stat().st_mtime, ref.stat().st_mtime, lib.stat().st_mtime):
        return out
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", str(REPO / "include"),
           str(src), str(ref), "-o", str(out), "-L", str(HERE), "-lcris_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../cris/pytorch_b200"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"selftest build failed:\n{res.stdout}\n{res.stderr}")
    return out


def build_halo_probe() -> Path:
    """tests/native/halo_probe: round-2 experiment (row-shifted UMMA descriptors on one swizzled TMA tile)."""
    src = REPO / "tests" / "native" / "halo_probe.cu"
    out = REPO / "tests" / "native" / "halo_probe"
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", str(REPO / "include"),
           str(src), "-o", str(out), "-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"halo_probe build failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(p)
    if "--selftest" in sys.argv:
        print(build_selftest())
    if "--halo-probe" in sys.argv:
        print(build_halo_probe())
