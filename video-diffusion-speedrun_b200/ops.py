"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers + stream."""
import ctypes

import torch

from . import lib as L


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_bf16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.bfloat16, (t.device, t.dtype)


def gemm(a, b, *, a_mn=False, b_mn=False, epilogue=L.EPI_STORE, out=None, out2=None, bias=None,
         aux=None, gate=None, rows_per_batch=0, splits=1, remap=None, M=None, N=None, K=None):
    """D[M,N] = A * B^T on tcgen05 (see include/vds_b200.h: vds_gemm).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True).
    2-D views with unit inner stride are accepted (leading dim = stride(0)).
    """
    _chk_bf16(a, b, bias, aux, gate)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    assert (b.shape[0] if b_mn else b.shape[1]) == K, (a.shape, b.shape, a_mn, b_mn)
    f32_out = epilogue in (L.EPI_ACCUM_F32, L.EPI_STORE_F32)
    if out is None and epilogue not in (L.EPI_ACCUM_F32,):
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    if epilogue in (L.EPI_BIAS_GELU, L.EPI_GATE_RES) and out2 is None:
        out2 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    args = L.GemmArgs()
    args.A, args.B = a.data_ptr(), b.data_ptr()
    args.lda, args.ldb = a.stride(0), b.stride(0)
    args.M, args.N, args.K = M, N, K
    args.a_mn, args.b_mn = int(a_mn), int(b_mn)
    args.epilogue, args.splits = epilogue, splits
    if out is not None:
        assert out.stride(-1) == 1 and out.dtype == (torch.float32 if f32_out else torch.bfloat16)
        args.C, args.ldc = out.data_ptr(), out.stride(-2)
    if out2 is not None:
        assert out2.stride(-1) == 1
        args.C2, args.ldc2 = out2.data_ptr(), out2.stride(-2)
    if bias is not None:
        args.bias = bias.data_ptr()
    if aux is not None:
        args.aux, args.ldaux = aux.data_ptr(), aux.stride(-2)
    if gate is not None:
        args.gate, args.gate_stride = gate.data_ptr(), gate.stride(0)
    args.rows_per_batch = rows_per_batch
    if remap is not None:
        args.remap_rows, args.remap_stride, args.remap_offset = remap
    L.check(L.lib().vds_gemm(ctypes.byref(args), _stream()), "vds_gemm")
    return (out, out2) if out2 is not None else out
