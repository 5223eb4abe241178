"""ctypes binding of ``libvds_b200.so`` (C ABI declared in ``include/vds_b200.h``).

There is deliberately NO fallback: if the shared object is missing, or a launch fails, we raise.
"""
import ctypes
import os
import subprocess

from . import PKG_DIR

CSRC = os.path.join(PKG_DIR, "csrc")
SO_PATH = os.path.join(CSRC, "libvds_b200.so")

i32, i64, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p
f32 = ctypes.c_float


class VdsError(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a into ``csrc/libvds_b200.so`` (nvcc, no GPU needed)."""
    r = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise VdsError("building libvds_b200.so failed")
    return SO_PATH


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", vp), ("B", vp), ("lda", i64), ("ldb", i64),
        ("M", i32), ("N", i32), ("K", i32), ("a_mn", i32), ("b_mn", i32),
        ("epilogue", i32), ("splits", i32),
        ("C", vp), ("ldc", i64), ("C2", vp), ("ldc2", i64),
        ("bias", vp), ("aux", vp), ("ldaux", i64),
        ("gate", vp), ("gate_stride", i64), ("rows_per_batch", i32),
        ("remap_rows", i32), ("remap_stride", i32), ("remap_offset", i32), ("tile_n", i32), ("cluster", i32),
        ("rope_tab", vp), ("v0", vp), ("ldv0", i64), ("lambda_", vp),
    ]


ERR_UNSUPPORTED = -3
EPI_STORE, EPI_ACCUM_F32, EPI_BIAS_GELU, EPI_GATE_RES, EPI_DGELU, EPI_STORE_F32, EPI_STORE_ROWDOT, EPI_QKV_ROPE = range(8)

fp = ctypes.POINTER(ctypes.c_float)

# name -> argtypes; must match include/vds_b200.h (checked by tests/test_abi.py)
_SIGNATURES = {
    "vds_gemm": [ctypes.POINTER(GemmArgs), vp],
    "vds_patchify": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "vds_unpatchify": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "vds_rope_rows": [vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp],
    "vds_rope_pack": [vp, vp, vp, i32, vp],
    "vds_timestep_embedding": [vp, vp, i32, i32, f32, vp],
    "vds_silu": [vp, vp, i64, vp],
    "vds_silu_bwd": [vp, vp, vp, i64, vp],
    "vds_rmsnorm_mod_fwd": [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, f32, vp],
    "vds_rmsnorm_mod_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, i32, i32, vp],
    "vds_gate_bwd": [vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, vp],
    "vds_qkv_post_fwd": [vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, vp],
    "vds_qkv_post_bwd": [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "vds_colsum": [vp, vp, i64, i32, i64, vp],
    "vds_batch_rowsum": [vp, vp, i32, i64, i32, i32, vp],
    "vds_cast_f32_bf16": [vp, vp, i64, f32, vp],
    "vds_cast_f32_bf16_2d": [vp, i64, vp, i64, i64, i32, f32, vp],
    "vds_accum_bf16_f32": [vp, vp, i64, i32, vp],
    "vds_attn_fwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i32, i32, i32, i32, i32, f32, vp],
    "vds_attn_bwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i64, vp, i64, vp, i64, vp, vp,
                     i64, i32, i32, i32, i32, i32, i32, f32, vp, i64, vp],
    "vds_debug_attn_bwd_trace": [vp],
    "vds_debug_attn_pair_mode": [i32],
    "vds_debug_gemm2_trace": [vp],
    "vds_loss_fwd_bwd": [vp, vp, vp, vp, vp, vp, i32, i64, f32, vp, vp],
    "vds_adamw": [vp, vp, vp, vp, vp, vp, vp, vp, i32, fp, fp, i32, f32, f32, f32, i32, f32, vp, vp],
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise VdsError(
                f"{SO_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the CUDA path)")
        L = ctypes.CDLL(SO_PATH)
        L.vds_last_error.restype = ctypes.c_char_p
        L.vds_abi_version.restype = ctypes.c_int
        L.vds_launch_count.restype = ctypes.c_int64
        L.vds_attn_bwd_tail_ws_bytes.restype = ctypes.c_int64
        L.vds_attn_bwd_tail_ws_bytes.argtypes = [i32, i32, i32]
        L.vds_attn_bwd_tail_plan.restype = ctypes.c_int      # returns a count, not an error code
        L.vds_attn_bwd_tail_plan.argtypes = [i32, i32, i32, ctypes.POINTER(ctypes.c_uint32), i32]
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise VdsError(f"{what} failed (rc={rc}): {lib().vds_last_error().decode()}")


def launch_count():
    return int(lib().vds_launch_count())
