"""tcgen05 GEMM vs torch fp32 matmul on the same bf16 inputs (all operand majors, tails, epilogues)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def _close(got, ref, tol=2e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, f"max err {err} vs scale {den}"


SHAPES = [(128, 128, 64), (256, 384, 512), (300, 128, 192), (16416, 512, 512), (2, 4608, 512),
          (1056, 1536, 512), (1024, 1024, 4096)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_kk_store(cuda_dev, M, N, K):
    import vds_b200
    from vds_b200 import ops
    a, b = _mk((M, K), cuda_dev, 1), _mk((N, K), cuda_dev, 2)
    bias = _mk((N,), cuda_dev, 3)
    out = ops.gemm(a, b, bias=bias)
    ref = a.float() @ b.float().t() + bias.float()
    _close(out, ref)
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(256, 384, 512), (300, 128, 192), (16416, 512, 1536)])
def test_gemm_dgrad_k_mn(cuda_dev, M, N, K):
    """dx[M,N] = dy[M,K] @ W[K,N]: A K-major, B MN-major."""
    from vds_b200 import ops
    dy, w = _mk((M, K), cuda_dev, 4), _mk((K, N), cuda_dev, 5)
    out = ops.gemm(dy, w, b_mn=True)
    _close(out, dy.float() @ w.float())


@pytest.mark.parametrize("M,N,K,splits", [(384, 256, 512, 1), (1536, 512, 16416, 4), (128, 128, 2, 1),
                                          (512, 4096, 1024, 3)])
def test_gemm_wgrad_mn_mn(cuda_dev, M, N, K, splits):
    """dW[M,N] += dy[K,M]^T @ x[K,N]: both MN-major, fp32 accumulate, split-K."""
    from vds_b200 import ops, lib
    dy, x = _mk((K, M), cuda_dev, 6), _mk((K, N), cuda_dev, 7)
    out = torch.ones((M, N), device=cuda_dev, dtype=torch.float32)
    ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=lib.EPI_ACCUM_F32, out=out, splits=splits)
    _close(out, 1.0 + dy.float().t() @ x.float())


def test_gemm_mn_k(cuda_dev):
    from vds_b200 import ops
    a, b = _mk((512, 384), cuda_dev, 8), _mk((256, 512), cuda_dev, 9)   # a: [K,M], b: [N,K]
    out = ops.gemm(a, b, a_mn=True)
    _close(out, a.float().t() @ b.float().t())


def test_gemm_bias_gelu(cuda_dev):
    from vds_b200 import ops, lib
    a, b, bias = _mk((528, 512), cuda_dev, 10), _mk((2048, 512), cuda_dev, 11, 0.05), _mk((2048,), cuda_dev, 12)
    pre, act = ops.gemm(a, b, bias=bias, epilogue=lib.EPI_BIAS_GELU)
    ref_pre = (a.float() @ b.float().t() + bias.float()).bfloat16()
    _close(pre, ref_pre)
    _close(act, torch.nn.functional.gelu(pre.float()), tol=1e-2)


def test_gemm_gate_res(cuda_dev):
    from vds_b200 import ops, lib
    Bt, Lr, h = 2, 272, 512
    a, w = _mk((Bt * Lr, h), cuda_dev, 13), _mk((h, h), cuda_dev, 14, 0.05)
    x, gate = _mk((Bt * Lr, h), cuda_dev, 15), _mk((Bt, 9 * h), cuda_dev, 16)
    g = gate[:, 2 * h:3 * h]
    lin, xo = ops.gemm(a, w, epilogue=lib.EPI_GATE_RES, aux=x, gate=g, rows_per_batch=Lr)
    ref_lin = (a.float() @ w.float().t()).bfloat16()
    _close(lin, ref_lin)
    ref = x + lin.view(Bt, Lr, h) .mul(g[:, None, :]).view(Bt * Lr, h)
    _close(xo, ref, tol=1e-2)


def test_gemm_dgelu(cuda_dev):
    from vds_b200 import ops, lib
    dy, w, h1 = _mk((528, 512), cuda_dev, 17), _mk((512, 2048), cuda_dev, 18, 0.05), _mk((528, 2048), cuda_dev, 19)
    out = ops.gemm(dy, w, b_mn=True, epilogue=lib.EPI_DGELU, aux=h1)
    hh = h1.float().requires_grad_(True)
    torch.nn.functional.gelu(hh).backward(dy.float() @ w.float())
    _close(out, hh.grad)


def test_gemm_remap_rows(cuda_dev):
    from vds_b200 import ops
    Bt, Np, h = 2, 256, 512
    a, w, bias = _mk((Bt * Np, 128), cuda_dev, 20), _mk((h, 128), cuda_dev, 21), _mk((h,), cuda_dev, 22)
    out = torch.zeros((Bt, Np + 16, h), device=cuda_dev, dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, out=out.view(-1, h), remap=(Np, Np + 16, 16))
    ref = (a.float() @ w.float().t() + bias.float()).view(Bt, Np, h)
    _close(out[:, 16:], ref)
    assert out[:, :16].abs().max().item() == 0


@pytest.mark.parametrize("tile_n", [0, 128])
def test_gemm_wide_tiles_epilogues(cuda_dev, tile_n):
    """128 x 256 tiles (auto) vs forced 128 x 128 on shapes large enough to pick the wide path."""
    from vds_b200 import ops, lib
    M, h = 4 * 4104, 512
    a, w1, b1 = _mk((M, h), cuda_dev, 30), _mk((4 * h, h), cuda_dev, 31, 0.05), _mk((4 * h,), cuda_dev, 32)
    pre, act = ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU, tile_n=tile_n)
    ref_pre = (a.float() @ w1.float().t() + b1.float()).bfloat16()
    _close(pre, ref_pre)
    _close(act, torch.nn.functional.gelu(pre.float()), tol=1e-2)
    w2, x, gate = _mk((h, 4 * h), cuda_dev, 33, 0.05), _mk((M, h), cuda_dev, 34), _mk((4, 9 * h), cuda_dev, 35)
    g = gate[:, 8 * h:]
    lin, xo = ops.gemm(act, w2, epilogue=lib.EPI_GATE_RES, aux=x, gate=g, rows_per_batch=4104, tile_n=tile_n)
    ref_lin = (act.float() @ w2.float().t()).bfloat16()
    _close(lin, ref_lin)
    _close(xo, x + (lin.view(4, 4104, h) * g[:, None, :]).view(M, h), tol=1e-2)
    dy = _mk((M, h), cuda_dev, 36)
    dh = ops.gemm(dy, w2, b_mn=True, epilogue=lib.EPI_DGELU, aux=pre, tile_n=tile_n)
    hh = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(hh).backward(dy.float() @ w2.float())
    _close(dh, hh.grad)


@pytest.mark.parametrize("cluster", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(304, 256, 512), (16416, 1536, 512), (1296, 512, 2048)])
def test_gemm_cluster_multicast_matches(cuda_dev, M, N, K, cluster):
    """2-CTA clusters with multicast B (auto) vs plain launch; odd numbers of m-tiles leave one CTA of a pair idle."""
    from vds_b200 import ops, lib
    a, b, bias = _mk((M, K), cuda_dev, 40), _mk((N, K), cuda_dev, 41), _mk((N,), cuda_dev, 42)
    _close(ops.gemm(a, b, bias=bias, cluster=cluster), a.float() @ b.float().t() + bias.float())
    dy, w = _mk((M, K), cuda_dev, 43), _mk((K, N), cuda_dev, 44)
    _close(ops.gemm(dy, w, b_mn=True, cluster=cluster), dy.float() @ w.float())
    g1, g2 = _mk((K, M), cuda_dev, 45), _mk((K, N), cuda_dev, 46)
    out = torch.zeros((M, N), device=cuda_dev, dtype=torch.float32)
    ops.gemm(g1, g2, a_mn=True, b_mn=True, epilogue=lib.EPI_ACCUM_F32, out=out, splits=2, cluster=cluster)
    _close(out, g1.float().t() @ g2.float())


def test_dgrad_rowdot_epilogue_matches_separate_kernels(cuda_dev):
    """VDS_EPI_STORE_ROWDOT: the dgrad GEMM that produces dO also emits delta[b, head, r] = <dO_head, O_head>
    (what attn_bwd_prep computes from the stored bf16 dO).  dX must be bit-identical to the plain GEMM."""
    from vds_b200 import ops
    torch.manual_seed(0)
    B, L, h, nh = 2, 8208, 512, 4
    M = B * L
    dy = torch.randn((M, h), device=cuda_dev).bfloat16()
    w = (torch.randn((h, h), device=cuda_dev) * 0.05).bfloat16()
    o = torch.randn((M, h), device=cuda_dev).bfloat16()
    ref_dx = ops.gemm(dy, w, b_mn=True)
    delta = torch.zeros((B, nh, L), device=cuda_dev, dtype=torch.float32)
    dx = ops.gemm_dgrad_rowdot(dy, w, o, delta, L)
    assert dx is not None, "2-CTA path expected for this shape"
    assert torch.equal(dx, ref_dx)
    ref_delta = (ref_dx.float() * o.float()).view(B, L, nh, 128).sum(-1).permute(0, 2, 1)
    assert torch.allclose(delta, ref_delta, rtol=1e-4, atol=1e-3), (delta - ref_delta).abs().max().item()
    # shapes without a 2-CTA tile path report "unsupported" (None) instead of computing something else
    small = ops.gemm_dgrad_rowdot(dy[:256], w, o[:256], torch.zeros((1, nh, 256), device=cuda_dev), 256)
    assert small is None


def test_gemm_xl_width_narrow_last_tile(cuda_dev):
    """DiT-XL width (h = 1152, 9 heads): N = 1152 / 3456 are not multiples of the 256-wide 2-CTA tile; the last tile is
    computed on TMA zero-fill and clipped.  Every epilogue of the step, at shapes large enough for the 2-CTA path."""
    from vds_b200 import ops, lib
    Bt, Lr, h = 4, 2064 * 2, 1152
    M = Bt * Lr
    a = _mk((M, h), cuda_dev, 50)
    # qkv: STORE with bias, N = 3456
    wq, bq = _mk((3 * h, h), cuda_dev, 51, 0.03), _mk((3 * h,), cuda_dev, 52)
    _close(ops.gemm(a, wq, bias=bq), a.float() @ wq.float().t() + bq.float())
    # attn_proj: GATE_RES, N = 1152
    wp, x, gate = _mk((h, h), cuda_dev, 53, 0.03), _mk((M, h), cuda_dev, 54), _mk((Bt, 9 * h), cuda_dev, 55)
    g = gate[:, 2 * h:3 * h]
    lin, xo = ops.gemm(a, wp, epilogue=lib.EPI_GATE_RES, aux=x, gate=g, rows_per_batch=Lr)
    ref_lin = (a.float() @ wp.float().t()).bfloat16()
    _close(lin, ref_lin)
    _close(xo, x + (lin.view(Bt, Lr, h) * g[:, None, :]).view(M, h), tol=1e-2)
    # dgrad of attn_proj with the attention delta (STORE_ROWDOT), N = 1152 = 9 heads
    dy, o = _mk((M, h), cuda_dev, 56), _mk((M, h), cuda_dev, 57)
    rowdot = torch.zeros((Bt, h // 128, Lr), device=cuda_dev, dtype=torch.float32)
    dx = ops.gemm_dgrad_rowdot(dy, wp, o, rowdot, Lr)
    assert dx is not None
    ref_dx = (dy.float() @ wp.float()).bfloat16()
    _close(dx, ref_dx)
    ref_dot = (ref_dx.float() * o.float()).view(Bt, Lr, h // 128, 128).sum(-1).permute(0, 2, 1)
    _close(rowdot, ref_dot, tol=2e-2)
    # mlp: BIAS_GELU N = 4608 (multiple of 256) then DGELU dgrad back to N = 4608 and the mlp.2 GATE_RES N = 1152
    w1, b1 = _mk((4 * h, h), cuda_dev, 58, 0.03), _mk((4 * h,), cuda_dev, 59)
    pre, act = ops.gemm(a, w1, bias=b1, epilogue=lib.EPI_BIAS_GELU)
    w2, b2 = _mk((h, 4 * h), cuda_dev, 60, 0.02), _mk((h,), cuda_dev, 61)
    lin2, xo2 = ops.gemm(act, w2, bias=b2, epilogue=lib.EPI_GATE_RES, aux=x, gate=g, rows_per_batch=Lr)
    ref_lin2 = (act.float() @ w2.float().t() + b2.float()).bfloat16()
    _close(lin2, ref_lin2)
    _close(xo2, x + (lin2.view(Bt, Lr, h) * g[:, None, :]).view(M, h), tol=1e-2)
    # wgrad of qkv: M = 3456, N = 1152, fp32 split-K accumulate
    dqkv = _mk((M, 3 * h), cuda_dev, 62)
    gw = torch.zeros((3 * h, h), device=cuda_dev, dtype=torch.float32)
    ops.gemm(dqkv, a, a_mn=True, b_mn=True, epilogue=lib.EPI_ACCUM_F32, out=gw, splits=4)
    _close(gw, dqkv.float().t() @ a.float())
    # dgrad of qkv: [M, 3456] x [3456, 1152] -> N = 1152
    _close(ops.gemm(dqkv, wq, b_mn=True), dqkv.float() @ wq.float())
