"""ctypes binding of libcris_b200.so (the C ABI declared in include/cris_b200.h).

The product path has NO fallback: if the library cannot be loaded, or a call fails, a RuntimeError is
raised (the reference's error convention is Python exceptions, e.g. model/layers.py:113-115).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libcris_b200.so"


class GemmArgs(C.Structure):
    """Mirror of `cris_gemm_args` (include/cris_b200.h)."""
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int64), ("strideA", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("strideB", C.c_int64),
        ("D", C.c_void_p), ("ldd", C.c_int64), ("strideD", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("batch", C.c_int32),
        ("a_mn", C.c_int32), ("b_mn", C.c_int32),
        ("d_fp32", C.c_int32), ("accumulate", C.c_int32),
        ("tap_mode", C.c_int32), ("taps", C.c_int32),
        ("tap_off", C.c_int32 * 9),
        ("b_tap_k", C.c_int32), ("b_tap_n", C.c_int32), ("d_tap_n", C.c_int32),
        ("splits", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64), ("strideR", C.c_int64), ("resid_fp32", C.c_int32),
        ("mask_hp", C.c_int32), ("mask_wp", C.c_int32),
        ("colstats", C.c_void_p),
        ("a_rows", C.c_int32), ("b_rows", C.c_int32),
        ("batch_inner", C.c_int32),
        ("strideA2", C.c_int64), ("strideB2", C.c_int64), ("strideD2", C.c_int64), ("strideR2", C.c_int64),
        ("d_col_stride", C.c_int32),
    ]


# signature mini-language: p = pointer, q = int64, i = int32, f = float, d = double, u = uint64
_T = {"p": C.c_void_p, "q": C.c_int64, "i": C.c_int32, "f": C.c_float, "d": C.c_double, "u": C.c_uint64}
_SIGS = {
    "cris_col_reduce": "ipqipqpqpqippppqiiiipip",
    "cris_bn_finalize_fwd": "piipdppffppppppp",
    "cris_stats_finalize_bwd": "piipppp",
    "cris_bn_reduce_partials": "piipp",
    "cris_bn_coeffs": "pdppffppppppiip",
    "cris_bn_bwd_reduce_masked": "pqpqpqppqiiipqpip",
    "cris_bn_apply": "pqpppqpqqiiiip",
    "cris_bn_bwd_apply": "pqpqpqpppppdpqpqiqiiiip",
    "cris_layernorm_fwd": "piqpppqipiqpqppqifp",
    "cris_layernorm_bwd": "piqpqpiqppppiqippqip",
    "cris_avgpool2_fwd": "pqpqiiiip",
    "cris_avgpool2_bwd": "pqpqiiiiip",
    "cris_upsample2x_fwd": "pqpqiiiip",
    "cris_upsample2x_bwd": "pqpqiiiiip",
    "cris_mul_bcast": "pqpqpqqiip",
    "cris_mul_bcast_bwd_s": "pqpqpiiip",
    "cris_padded_to_tokens": "pqpqpiqiiiip",
    "cris_tokens_to_padded": "piqpqiiiip",
    "cris_coord_fill": "pqiiiip",
    "cris_stem_im2col": "ppiiip",
    "cris_stem_conv1_fwd": "pppqiiiip",
    "cris_stem_conv1_wgrad": "ppqpiiiip",
    "cris_softmax_fwd": "pppqqiiiipifupp",
    "cris_softmax_bwd": "ppqqiiifupp",
    "cris_embed_fwd": "ppppiiip",
    "cris_embed_bwd": "ppppiiip",
    "cris_eot_gather": "ppiqpqiiip",
    "cris_eot_scatter": "ppiqpiqiiip",
    "cris_elementwise": "ipiqpiqpiqqifupp",
    "cris_pack_conv_weight": "ppiiiip",
    "cris_unpack_conv_wgrad": "ppiiiip",
    "cris_pack_multi": "piqp",
    "cris_bn_coeffs_multi": "piifp",
    "cris_pack_conv_weight_scaled": "pppiiiip",
    "cris_pack_matrix_scaled": "pppqiip",
    "cris_pack_matrix": "ppqiip",
    "cris_batch_reduce": "piqpqiiiip",
    "cris_small_matmul": "pppiiiiip",
    "cris_dynconv_bce_fwd": "pqpqpiippppfiiiip",
    "cris_dynconv_bce_bwd": "pqpqpppppqpqiiiip",
    "cris_adam_step": "piqddddddppp",
    "cris_postproc_upsample": "ppiiiiip",
    "cris_postproc_warp_iou": "piiipppfpqp",
    "cris_feeder_letterbox": "pppiiippppp" "p",
    "cris_attention_fwd": "pqpqpqpqpiiiiffupp",
    "cris_attention_bwd": "pqpqpqpqpq" "pppqpqpq" "iiii" "ffu" "pp",
    "cris_conv3x3_halo": "pqpqipqpiiiiip",
    "cris_pack_conv_weight_dgrad": "ppiiip",
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cris.pytorch_b200.build` "
                "(cris.pytorch_b200 has no CPU / eager fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.cris_last_error.restype = C.c_char_p
        L.cris_launch_count.restype = C.c_uint64
        L.cris_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
        L.cris_gemm.restype = C.c_int
        L.cris_gemm_plan.argtypes = [C.POINTER(GemmArgs), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.cris_add_launch_count.argtypes = [C.c_uint64]
        L.cris_add_launch_count.restype = None
        L.cris_peer_buffer_bytes.restype = C.c_size_t
        L.cris_peer_buffer_create.argtypes = [C.POINTER(C.c_void_p), C.c_char_p]
        L.cris_peer_buffer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.cris_peer_buffer_close.argtypes = [C.c_void_p, C.c_int]
        L.cris_peer_allreduce_f32.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_double, C.c_void_p]
        L.cris_peer_bn_sync_fwd.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_double, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                            C.c_void_p]
        L.cris_peer_bn_sync_bwd.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = [_T[ch] for ch in sig]
            fn.restype = C.c_int
        if L.cris_gemm_args_size() != C.sizeof(GemmArgs) or \
                L.cris_gemm_args_last_offset() != GemmArgs.d_col_stride.offset:
            raise RuntimeError("cris_gemm_args layout mismatch between include/cris_b200.h and _lib.GemmArgs")
        _lib = L
    return _lib


def exported_symbols():
    """Every entry point include/cris_b200.h declares (used by the CPU 'library loads' test)."""
    return ["cris_last_error", "cris_abi_version", "cris_device_check",
            "cris_launch_count", "cris_add_launch_count", "cris_debug_set_trace", "cris_gemm", "cris_gemm_plan", "cris_gemm_args_size", "cris_gemm_args_last_offset",
            "cris_peer_buffer_bytes", "cris_peer_buffer_create", "cris_peer_buffer_open", "cris_peer_buffer_close",
            "cris_peer_allreduce_f32", "cris_peer_bn_sync_fwd", "cris_peer_bn_sync_bwd", "cris_postproc_sample_bytes", "cris_feeder_sample_bytes", "cris_pack_entry_bytes", "cris_pack_chunk_elems", "cris_bn_eval_entry_bytes", "cris_adam_table_entry_bytes", "cris_adam_chunk_elems", *_SIGS.keys()]


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"libcris_b200 {what} failed: {lib().cris_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---- live per-launch timing (bench.py's per-class roofline table; eager launches only) ------------------------
_prof = None       # None = off; else a list of [entry point, class tag, flop, event0, event1]
prof_tag = None    # set by the engine around a launch: (class, algorithmic flop) of the next call


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> [(entry point, class, flop, milliseconds)] of every launch since profile_begin()."""
    global _prof
    rows, _prof = _prof or [], None
    torch.cuda.synchronize()
    return [(n, c, f, e0.elapsed_time(e1)) for n, c, f, e0, e1 in rows]


def _timed(name, fn):
    global prof_tag
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    cls, flop = prof_tag if prof_tag is not None else (None, 0.0)
    prof_tag = None
    _prof.append([name, cls, flop, e0, e1])
    return rc


def call(name: str, *args):
    """Invoke an entry point on the current CUDA stream (appended as the last argument)."""
    if _prof is not None:
        rc = _timed(name, lambda: getattr(lib(), name)(*args, stream_ptr()))
    else:
        rc = getattr(lib(), name)(*args, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"libcris_b200 {name} failed: {lib().cris_last_error().decode()}")


def gemm(args: GemmArgs):
    if _prof is not None:
        rc = _timed("cris_gemm", lambda: lib().cris_gemm(C.byref(args), stream_ptr()))
    else:
        rc = lib().cris_gemm(C.byref(args), stream_ptr())
    if rc != 0:
        raise RuntimeError(f"libcris_b200 cris_gemm failed: {lib().cris_last_error().decode()}")


_checked_devices = set()


def device_check():
    """Raise unless the current device is sm_100; cached per device (cudaGetDeviceProperties is slow and jittery)."""
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
    if dev in _checked_devices:
        return
    rc = lib().cris_device_check()
    if rc != 0:
        raise RuntimeError(f"libcris_b200: {lib().cris_last_error().decode()}")
    _checked_devices.add(dev)


def launch_count() -> int:
    return int(lib().cris_launch_count())
