"""ctypes binding of libsct_b200.so (the C ABI declared in include/sct_b200.h).

PyTorch is only the owner of device memory and streams here: every call passes raw device pointers
(`tensor.data_ptr()`) and the current CUDA stream handle.  There is no fallback of any kind: if the
library is missing or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SCT_B200_LIB", _PKG / "libsct_b200.so"))

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_u64 = C.c_uint64
_f = C.c_float

# name -> argtypes, mirrors include/sct_b200.h one to one
SIGNATURES = {
    "sct_embed_ln_pe_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _f, _f, _u64, _u64, _p, _p],
    "sct_embed_ln_pe_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _f, _u64, _u64, _p, _p],
    "sct_add_dropout_ln_fwd": [_p, _p, _f, _p, _p, _p, _p, _p, _p, _i64, _i64, _f, _u64, _u64, _p, _p],
    "sct_add_dropout_ln_bwd": [_p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _i64, _i64, _f, _u64, _u64, _p, _p],
    "sct_ln_act_fwd": [_p, _p, _p, _p, _p, _i64, _i64, _f, _u64, _u64, _p, _p],
    "sct_ln_act_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _f, _u64, _u64, _p, _p],
    "sct_gelu_dropout_fwd": [_p, _p, _i64, _f, _u64, _u64, _p, _p],
    "sct_gelu_dropout_bwd": [_p, _p, _p, _i64, _f, _u64, _u64, _p, _p],
    "sct_colsum_bf16": [_p, _i64, _p, _i64, _i64, _f, _p],
    "sct_cast_scale": [_p, _p, _i64, _p, _i64, _i64, _i64, _i64, _f, _p],
    "sct_seq_mean_fwd": [_p, _p, _p, _i64, _i64, _i64, _p],
    "sct_seq_mean_bwd": [_p, _p, _p, _i64, _i64, _i64, _p],
    "sct_gemm_bf16_nt": [_p, _i64, _p, _i64, _p, _i64, _p, _f, _i64, _i64, _i64, _i32, _p],
    "sct_gemm_bf16_nn": [_p, _i64, _p, _i64, _p, _i64, _p, _f, _i64, _i64, _i64, _i32, _p],
    "sct_gemm_bf16_tn": [_p, _i64, _p, _i64, _p, _i64, _f, _i64, _i64, _i64, _i32, _p],
    "sct_gemm_bf16_nt_gelu": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _f, _u64, _u64, _p, _p],
    "sct_gemm_bf16_nn_mul": [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _p],
    "sct_gemm_bf16_tn_colsum": [_p, _i64, _p, _i64, _p, _i64, _p, _f, _i64, _i64, _i64, _i32, _p],
    "sct_attn_fwd": [_p, _i64, _p, _p, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _f, _f,
                     _u64, _u64, _p, _p],
    "sct_attn_fwd_strided": [_p, _i64, _p, _p, _i64, _i64, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i32, _f,
                             _f, _u64, _u64, _p, _p],
    "sct_attn_bwd": [_p, _i64, _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _i64,
                     _i64, _i64, _i64, _i32, _f, _f, _u64, _u64, _p, _p],
    "sct_attn_bwd_ws": [_p, _i64, _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _i64,
                        _i64, _i64, _i64, _i32, _f, _f, _u64, _u64, _p, _p, _i64, _p],
    "sct_attn_bwd_workspace_bytes": [_i64, _i64, _i64, _i64],
    "sct_ce_rows": [_p, _p, _p, _p, _i64, _i64, _i64, _f, _i32, _p],
    "sct_sample_rows": [_p, _i64, _i64, _i64, _f, _i32, _f, _i32, _p, _u64, _u64, _p, _p, _p],
    "sct_small_linear_fwd": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "sct_small_linear_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "sct_gan_loss_fwd": [_p, _i64, _p, _p, _p],
    "sct_gan_loss_bwd": [_p, _i64, _p, _p, _p, _p, _p],
    "sct_clip_adamw_step": [_p, _i32, _p, _i32, _p, _p, _p, _f, _f, _f, _f, _f, _f, _p],
}
NOARG = {"sct_opt_chunk_elems": _i32, "sct_version": _i32, "sct_device_check": _i32, "sct_debug_timeouts": _i32, "sct_last_error": C.c_char_p}

RESTYPE = {"sct_attn_bwd_workspace_bytes": _i64}  # everything else returns an int32 status
_lib = None


def exported_symbols() -> list[str]:
    return sorted(list(SIGNATURES) + list(NOARG))


def load() -> C.CDLL:
    """Loads the library (once).  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m sct_gan_b200.build` "
            "(or __graft_entry__.build()); sct_gan_b200 has no CPU or PyTorch fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = RESTYPE.get(name, _i32)
    for name, res in NOARG.items():
        fn = getattr(lib, name)
        fn.argtypes = []
        fn.restype = res
    _lib = lib
    return lib


def last_error() -> str:
    return load().sct_last_error().decode("utf-8", "replace")


# kernels launched per C-ABI call (sct_attn_bwd = D-vector + dK/dV + dQ, sct_seq_mean_fwd = memset + reduce, ...)
KERNELS_PER_CALL = {"sct_clip_adamw_step": 4, "sct_attn_bwd": 3, "sct_attn_bwd_ws": 3, "sct_small_linear_bwd": 2, "sct_seq_mean_fwd": 2}


class Stats:
    """Launch accounting + optional per-call CUDA-event timing (bench.py's roofline leg).

    `launches` counts device kernels launched through the C ABI.  With `events` set to a list, every call
    whose name is in `timed` (None = every call) is bracketed by CUDA events recorded on the launching (current torch) stream
    and appended as (name, work, start_event, end_event), `work` being the call's algorithmic flops/bytes
    supplied by the wrapper through `annotate`."""
    launches = 0
    events = None
    timed = ()
    _work = 0.0

    @classmethod
    def annotate(cls, work: float):
        cls._work = work


# entry points that run the same kernel are accounted under one name
STAT_NAME = {"sct_gemm_bf16_tn_colsum": "sct_gemm_bf16_tn", "sct_attn_bwd_ws": "sct_attn_bwd"}


def call(name: str, *args) -> None:
    fn = getattr(load(), name)
    ev = Stats.events
    name = STAT_NAME.get(name, name)
    if ev is not None and (Stats.timed is None or name in Stats.timed):
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        ev.append((name, Stats._work, e0, e1))
    else:
        rc = fn(*args)
    Stats._work = 0.0  # a call that does not annotate must not inherit the previous call's work
    Stats.launches += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {last_error()}")
