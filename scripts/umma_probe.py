#!/usr/bin/env python
"""Pins the shared-memory / tensor-memory layouts the 2-CTA attention backward depends on, on the hardware
(driver of scripts/umma_probe.cu; tuning aid, not product code).

Every case builds the operand images for ONE hypothesis, runs a few tcgen05.mma (kind::f16, bf16 in, fp32 out), reads
the whole accumulator region back (all 128 lanes x N columns of each CTA) and
  * checks the dump against the exact product (integer-coded operands), under the TMEM layout hypothesis of the case, or
  * with --decode, prints which (m, n) every (lane, column) holds (operands coded so that D[m, n] = (m+1) + 512 (n+1)).

  python scripts/umma_probe.py            # run every case, each in its own process (a bad descriptor may fault)
  python scripts/umma_probe.py --case s_t_cg2
"""
import argparse
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "scripts", "_build", "libumma_probe.so")


class ProbeArgs(ctypes.Structure):
    _fields_ = [("a_img", ctypes.c_void_p), ("b_img", ctypes.c_void_p), ("a_tmem", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("cycles", ctypes.c_void_p), ("a_desc", ctypes.c_uint64),
                ("b_desc", ctypes.c_uint64), ("a_bytes", ctypes.c_int), ("b_bytes", ctypes.c_int),
                ("a_tmem_cols", ctypes.c_int), ("a_step", ctypes.c_int), ("b_step", ctypes.c_int),
                ("idesc", ctypes.c_uint32), ("ksteps", ctypes.c_int), ("d_cols", ctypes.c_int), ("reps", ctypes.c_int),
                ("remote_b", ctypes.c_int), ("ts", ctypes.c_int), ("d_alt", ctypes.c_int)]


def bf16_bits(x):
    """fp32 array -> uint16 bf16 bit patterns (values are chosen exactly representable)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    return (u >> 16).astype(np.uint16)


def desc(lbo, sbo, swizzle):
    sw = {"none": 0, "128": 2, "64": 4, "32": 6}[swizzle]
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | sw << 61


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


# ---------------------------------------------------------------------------- shared-memory images
def img_kmajor_sw128(X):
    """X [rows, K] (K multiple of 64): K-major, 128-byte swizzle; 64-wide K groups rows*128 bytes apart."""
    rows, K = X.shape
    out = np.zeros(rows * K * 2, dtype=np.uint8).view(np.uint16)
    bits = bf16_bits(X)
    for g in range(K // 64):
        for r in range(rows):
            for c in range(8):
                dst = (g * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4)) // 2
                out[dst:dst + 8] = bits[r, g * 64 + c * 8: g * 64 + c * 8 + 8]
    return out.view(np.uint8)


def img_mnmajor_sw128(X):
    """X [K rows, MN] (MN multiple of 64): MN contiguous, 128-byte swizzle; 64-wide MN groups K*128 bytes apart."""
    K, MN = X.shape
    out = np.zeros(K * MN * 2, dtype=np.uint8).view(np.uint16)
    bits = bf16_bits(X)
    for g in range(MN // 64):
        for k in range(K):
            for c in range(8):
                dst = (g * K * 128 + k * 128 + ((c ^ (k & 7)) << 4)) // 2
                out[dst:dst + 8] = bits[k, g * 64 + c * 8: g * 64 + c * 8 + 8]
    return out.view(np.uint8)


def img_mnmajor_sw64(X):
    """X [K rows, 32]: MN contiguous (64-byte rows), 64-byte swizzle (16-byte chunk index ^= (row >> 1) & 3)."""
    K, MN = X.shape
    assert MN == 32
    out = np.zeros(K * 64, dtype=np.uint8).view(np.uint16)
    bits = bf16_bits(X)
    for k in range(K):
        for c in range(4):
            dst = (k * 64 + ((c ^ ((k >> 1) & 3)) << 4)) // 2
            out[dst:dst + 8] = bits[k, c * 8: c * 8 + 8]
    return out.view(np.uint8)


def img_mnmajor_noswz(X, mn_stride, k_stride):
    """X [K rows, MN]: core matrices of 8 k-rows x 8 MN elements (16 B per row, 128 B per core matrix, k-row j at
    +16 j); core matrices mn_stride bytes apart along MN and k_stride bytes apart along K."""
    K, MN = X.shape
    size = max((MN // 8 - 1) * mn_stride + (K // 8 - 1) * k_stride + 128, 128)
    out = np.zeros(size, dtype=np.uint8).view(np.uint16)
    bits = bf16_bits(X)
    for k in range(K):
        for c in range(MN // 8):
            dst = (c * mn_stride + (k // 8) * k_stride + (k % 8) * 16) // 2
            out[dst:dst + 8] = bits[k, c * 8: c * 8 + 8]
    return out.view(np.uint8)


def img_kmajor_noswz16(X):
    """X [rows, 16]: the K = 16 no-swizzle tile the statistics k-step uses: 8-row x 16-byte core matrices, the two
    8-column halves 128 B apart (LBO), 8-row groups 256 B apart (SBO)."""
    rows, K = X.shape
    assert K == 16
    out = np.zeros(rows // 8 * 256, dtype=np.uint8).view(np.uint16)
    bits = bf16_bits(X)
    for r in range(rows):
        for half in range(2):
            dst = ((r >> 3) * 256 + half * 128 + (r & 7) * 16) // 2
            out[dst:dst + 8] = bits[r, half * 8: half * 8 + 8]
    return out.view(np.uint8)


def pad16(b):
    n = (len(b) + 15) // 16 * 16
    o = np.zeros(n, dtype=np.uint8)
    o[:len(b)] = b
    return o


def run_probe(a_imgs, b_imgs, a_desc, b_desc, a_step, b_step, idsc, ksteps, d_cols, cg, remote_b=0, a_tmem=None,
              reps=1, nclusters=1, d_alt=1):
    import torch
    lib = ctypes.CDLL(SO)
    lib.umma_probe.argtypes = [ctypes.POINTER(ProbeArgs), ctypes.c_int, ctypes.c_int]
    ncta = cg
    a_imgs = [pad16(x) for x in a_imgs]
    b_imgs = [pad16(x) for x in b_imgs]
    ab, bb = max(len(x) for x in a_imgs), max(len(x) for x in b_imgs)
    A = np.zeros((ncta, ab), dtype=np.uint8)
    Bm = np.zeros((ncta, bb), dtype=np.uint8)
    for r in range(ncta):
        A[r, :len(a_imgs[r])] = a_imgs[r]
        Bm[r, :len(b_imgs[r])] = b_imgs[r]
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda()
    out = torch.zeros((ncta, 128, d_cols), device="cuda", dtype=torch.float32)
    cyc = torch.zeros(2, device="cuda", dtype=torch.int64)
    p = ProbeArgs()
    p.a_img, p.b_img, p.out, p.cycles = dA.data_ptr(), dB.data_ptr(), out.data_ptr(), cyc.data_ptr()
    p.a_desc, p.b_desc, p.a_bytes, p.b_bytes = a_desc, b_desc, ab, bb
    p.a_step, p.b_step, p.idesc, p.ksteps, p.d_cols, p.reps = a_step, b_step, idsc, ksteps, d_cols, reps
    p.remote_b, p.ts, p.d_alt = remote_b, 0, d_alt
    keep = None
    if a_tmem is not None:
        keep = torch.from_numpy(np.ascontiguousarray(a_tmem.astype(np.uint32)).view(np.int32)).cuda()
        p.a_tmem, p.a_tmem_cols, p.ts = keep.data_ptr(), a_tmem.shape[-1], 1
    rc = lib.umma_probe(ctypes.byref(p), cg, nclusters)
    assert rc == 0, rc
    return out.cpu().numpy(), cyc.cpu().numpy()


def coded(M, N, K):
    """A [M, K], B [N, K] with D[m, n] = (m + 1) + 512 (n + 1) from k = 0, 1 only (exact in bf16 / fp32)."""
    A = np.zeros((M, K), dtype=np.float32)
    B = np.zeros((N, K), dtype=np.float32)
    A[:, 0] = np.arange(1, M + 1)
    A[:, 1] = 1
    B[:, 0] = 1
    B[:, 1] = 512.0 * np.arange(1, N + 1)
    return A, B


def rand_ints(M, N, K, seed=0):
    g = np.random.default_rng(seed)
    return g.integers(-4, 5, (M, K)).astype(np.float32), g.integers(-4, 5, (N, K)).astype(np.float32)


def decode(out, tag):
    """Print which (m, n) each (lane, column) holds for coded() operands."""
    ncta, lanes, cols = out.shape
    for r in range(ncta):
        d = out[r]
        ok = np.isfinite(d) & (d > 0)
        m = np.where(ok, np.mod(d, 512) - 1, -1).astype(int)
        n = np.where(ok, d // 512 - 1, -1).astype(int)
        print(f"[{tag}] CTA {r}: lanes with data: {sorted(set(np.where(ok.any(1))[0].tolist()))[:4]}..{int(np.where(ok.any(1))[0].max()) if ok.any() else -1} "
              f"(count {int(ok.any(1).sum())}); columns with data: {int(ok.any(0).sum())}")
        for lane in (0, 1, 15, 16, 31, 32, 63, 64, 65, 95, 96, 127):
            row = [(int(m[lane, c]), int(n[lane, c])) for c in (0, 1, 15, 16, 31, 32, 33, 63) if c < cols]
            print(f"   lane {lane:3d}: (m,n) at cols 0,1,15,16,31,32,33,63 = {row}")


def check(out, exp_fn, tag):
    """exp_fn(r, lane, col) -> expected value or None (don't care)."""
    ncta, lanes, cols = out.shape
    bad = 0
    for r in range(ncta):
        for lane in range(lanes):
            for c in range(cols):
                e = exp_fn(r, lane, c)
                if e is None:
                    continue
                if not (out[r, lane, c] == e):
                    if bad < 5:
                        print(f"   [{tag}] mismatch CTA {r} lane {lane} col {c}: got {out[r, lane, c]} expected {e}")
                    bad += 1
    print(f"[{tag}] {'OK' if bad == 0 else 'FAIL (%d mismatches)' % bad}")
    return bad == 0


# ---------------------------------------------------------------------------- cases
def case_sanity_cg1(args):
    """cta_group::1 M128 N64 K64, K-major SW128 A and B: the configuration the shipped kernels use."""
    A, B = rand_ints(128, 64, 64)
    out, _ = run_probe([img_kmajor_sw128(A)], [img_kmajor_sw128(B)], desc(16, 1024, "128"), desc(16, 1024, "128"), 32, 32,
                       idesc(128, 64, 0, 0), 4, 64, 1)
    D = A @ B.T
    check(out, lambda r, l, c: D[l, c], "sanity cg1 M128 N64 K-major")


def case_s_t_cg2(args):
    """S^T / dP^T: cta_group::2 M256 N64, A = own 128 rows K-major SW128, B = 32 rows per CTA K-major SW128."""
    A, B = (coded(256, 64, 64) if args.decode else rand_ints(256, 64, 64, 1))
    out, cyc = run_probe([img_kmajor_sw128(A[:128]), img_kmajor_sw128(A[128:])],
                         [img_kmajor_sw128(B[:32]), img_kmajor_sw128(B[32:])],
                         desc(16, 1024, "128"), desc(16, 1024, "128"), 32, 32, idesc(256, 64, 0, 0), 4, 64, 2)
    if args.decode:
        return decode(out, "s_t_cg2")
    D = A @ B.T
    check(out, lambda r, l, c: D[r * 128 + l, c], "S^T cg2 M256 N64: lane = own row, col = n (B halves concatenated)")


def case_stat_cg2(args):
    """statistics k-step: cta_group::2 M256 N64 K16, no-swizzle K-major tiles (A 128 rows, B 32 rows per CTA)."""
    A, B = rand_ints(256, 64, 16, 2)
    out, _ = run_probe([img_kmajor_noswz16(A[:128]), img_kmajor_noswz16(A[128:])],
                       [img_kmajor_noswz16(B[:32]), img_kmajor_noswz16(B[32:])],
                       desc(128, 256, "none"), desc(128, 256, "none"), 0, 0, idesc(256, 64, 0, 0), 1, 64, 2)
    D = A @ B.T
    check(out, lambda r, l, c: D[r * 128 + l, c], "stats k-step cg2 no-swizzle K16")


def case_dv_cg2(args):
    """dV / dK: cta_group::2 M256 N128 K64, A (bf16 pairs) from TMEM, B MN-major SW128, one 64-wide atom per CTA."""
    A, B = rand_ints(256, 128, 64, 3)          # B [N, K]
    at = np.zeros((2, 128, 32), dtype=np.uint32)
    bits = bf16_bits(A).astype(np.uint32)
    for r in range(2):
        for k2 in range(32):
            at[r, :, k2] = bits[r * 128:(r + 1) * 128, 2 * k2] | (bits[r * 128:(r + 1) * 128, 2 * k2 + 1] << 16)
    Bt = B.T                                   # [K, N]
    out, _ = run_probe([np.zeros(16, np.uint8)] * 2, [img_mnmajor_sw128(Bt[:, :64]), img_mnmajor_sw128(Bt[:, 64:])],
                       0, desc(64 * 128, 1024, "128"), 8, 2048, idesc(256, 128, 0, 1), 4, 128, 2, a_tmem=at)
    D = A @ B.T
    check(out, lambda r, l, c: D[r * 128 + l, c], "dV cg2 M256 N128 TS, B MN-major N-split 64|64")


def _dq_operands(args, seed):
    A, B = (coded(128, 64, 32) if args.decode else rand_ints(128, 64, 32, seed))   # A [M = d, K = kv], B [N = q, K]
    return A, B


def _dq_report(args, out, A, B, tag):
    if args.decode:
        return decode(out, tag)
    D = A @ B.T
    # hypothesis 1: CTA r holds rows m = 64 r + (lane % 64)?? — report which simple layouts match
    hyps = {
        "lane<64: m=64r+lane, col=n": lambda r, l, c: D[64 * r + l, c] if l < 64 else None,
        "m=64r+(lane%32)+32*(lane//64)... cols split": lambda r, l, c: None,
    }
    ok = check(out, hyps["lane<64: m=64r+lane, col=n"], tag + " [lanes 0-63 = rows, col = n]")
    if not ok:
        # layout with 128 lanes x N/2 columns: lane = 32*(2*(n//32)+ ... ) try: lanes 0-31 rows 0-31, 32-63 rows 32-63 for
        # cols n<32; lanes 64-127 the same rows for n >= 32 in columns 0..31
        def h2(r, l, c):
            if c >= 32:
                return None
            return D[64 * r + (l % 64), c + 32 * (l // 64)]
        check(out, h2, tag + " [lane%64 = row, lane//64 = column half, 32 cols]")


def case_dq_sw64(args):
    """dQ^T: cta_group::2 M128 N64 K32: A MN-major SW128 (64 rows of M per CTA), B MN-major 64-byte swizzle (32 per CTA)."""
    A, B = _dq_operands(args, 4)
    At, Bt = A.T, B.T    # [K, M], [K, N]
    out, _ = run_probe([img_mnmajor_sw128(At[:, :64]), img_mnmajor_sw128(At[:, 64:])],
                       [img_mnmajor_sw64(Bt[:, :32]), img_mnmajor_sw64(Bt[:, 32:])],
                       desc(32 * 128, 1024, "128"), desc(32 * 64, 512, "64"), 2048, 1024, idesc(128, 64, 1, 1), 2, 64, 2,
                       remote_b=args.remote)
    _dq_report(args, out, A, B, "dq sw64" + (" remote" if args.remote else ""))


def case_dq_noswz_a(args):
    """dQ^T, B MN-major no swizzle, descriptor (LBO, SBO) = (stride along MN, stride along K)."""
    A, B = _dq_operands(args, 5)
    At, Bt = A.T, B.T
    mn_stride, k_stride = 128, 512      # 4 core matrices along N (32 elements), then the next 8 k-rows
    out, _ = run_probe([img_mnmajor_sw128(At[:, :64]), img_mnmajor_sw128(At[:, 64:])],
                       [img_mnmajor_noswz(Bt[:, :32], mn_stride, k_stride), img_mnmajor_noswz(Bt[:, 32:], mn_stride, k_stride)],
                       desc(32 * 128, 1024, "128"), desc(mn_stride, k_stride, "none"), 2048, 2 * k_stride,
                       idesc(128, 64, 1, 1), 2, 64, 2)
    _dq_report(args, out, A, B, "dq noswz LBO=MN SBO=K")


def case_dq_noswz_b(args):
    """dQ^T, B MN-major no swizzle, descriptor (LBO, SBO) = (stride along K, stride along MN)."""
    A, B = _dq_operands(args, 6)
    At, Bt = A.T, B.T
    mn_stride, k_stride = 128, 512
    out, _ = run_probe([img_mnmajor_sw128(At[:, :64]), img_mnmajor_sw128(At[:, 64:])],
                       [img_mnmajor_noswz(Bt[:, :32], mn_stride, k_stride), img_mnmajor_noswz(Bt[:, 32:], mn_stride, k_stride)],
                       desc(32 * 128, 1024, "128"), desc(k_stride, mn_stride, "none"), 2048, 2 * k_stride,
                       idesc(128, 64, 1, 1), 2, 64, 2)
    _dq_report(args, out, A, B, "dq noswz LBO=K SBO=MN")


def case_dq_sw128half(args):
    """dQ^T, B MN-major SW128 atoms of which only the first 32 MN elements (64 B of every 128-byte row) are used."""
    A, B = _dq_operands(args, 7)
    At, Bt = A.T, B.T
    pad = lambda X: np.concatenate([X, np.zeros_like(X)], 1)   # noqa: E731  [K, 64]
    out, _ = run_probe([img_mnmajor_sw128(At[:, :64]), img_mnmajor_sw128(At[:, 64:])],
                       [img_mnmajor_sw128(pad(Bt[:, :32])), img_mnmajor_sw128(pad(Bt[:, 32:]))],
                       desc(32 * 128, 1024, "128"), desc(32 * 128, 1024, "128"), 2048, 2048, idesc(128, 64, 1, 1), 2, 64, 2)
    _dq_report(args, out, A, B, "dq sw128 (half-used atom)")


def case_timing(args):
    """cycles per MMA by shape, cta_group and accumulator rotation (all SMs busy), operands zero, 64-MMA series."""
    z = np.zeros(64 * 1024, np.uint8)
    K, MN = desc(16, 1024, "128"), desc(4096, 1024, "128")
    at = np.zeros((2, 128, 64), dtype=np.uint32)
    rows = [
        # name, cg, M, N, a_mn, b_mn, a_desc, b_desc, ts
        ("SS K-major", 1, 128, 64, 0, 0, K, K, 0), ("SS K-major", 1, 128, 128, 0, 0, K, K, 0),
        ("SS K-major", 1, 128, 256, 0, 0, K, K, 0),
        ("SS K-major", 2, 256, 64, 0, 0, K, K, 0), ("SS K-major", 2, 256, 128, 0, 0, K, K, 0),
        ("SS K-major", 2, 256, 256, 0, 0, K, K, 0),
        ("SS MN-major (B sw64)", 2, 128, 64, 1, 1, MN, desc(2048, 512, "64"), 0),
        ("TS, B MN-major", 1, 128, 128, 0, 1, 0, MN, 1), ("TS, B MN-major", 2, 256, 128, 0, 1, 0, MN, 1),
        ("TS, B K-major", 2, 256, 64, 0, 0, 0, K, 1),
    ]
    for name, cg, M, N, a_mn, b_mn, ad, bd, ts in rows:
        for alt in (1, 2):
            if alt * N > 384:
                continue
            imgs = [z] * cg
            _, cyc = run_probe(imgs, imgs, ad, bd, 0, 0, idesc(M, N, a_mn, b_mn), 64, 32, cg, reps=4,
                               nclusters=148 // cg, a_tmem=at[:cg] if ts else None, d_alt=alt)
            floor = max(M // cg, 128) * N / 256
            print(f"[timing] cta_group::{cg} M{M} N{N} {name}, {alt} accumulator(s): {cyc[1] / 64:.1f} cyc/MMA "
                  f"(floor {floor:.0f})", flush=True)


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="")
    ap.add_argument("--decode", action="store_true")
    ap.add_argument("--remote", type=int, default=0)
    a = ap.parse_args()
    if a.case:
        CASES[a.case](a)
        return
    runs = [(c, []) for c in CASES if c != "timing"]
    runs += [("s_t_cg2", ["--decode"]), ("dq_sw64", ["--decode"]), ("dq_noswz_a", ["--decode"]), ("dq_noswz_b", ["--decode"]),
             ("dq_sw128half", ["--decode"]), ("dq_sw64", ["--remote", "1"]), ("timing", [])]
    for c, extra in runs:
        print(f"===== {c} {' '.join(extra)}", flush=True)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c] + extra, capture_output=True, text=True,
                           timeout=300)
        print(r.stdout[-4000:])
        if r.returncode != 0:
            print("   rc", r.returncode, r.stderr[-1500:])


if __name__ == "__main__":
    main()
