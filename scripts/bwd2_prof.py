"""Accumulated wait / phase cycles of cluster 0 of the CTA-pair attention backward (library built with -DVDS_B2_PROF):
  cd video-diffusion-speedrun_b200/csrc && touch attention_bwd2.cu && make NVCCFLAGS_EXTRA=-DVDS_B2_PROF"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vds_b200
from vds_b200 import ops, lib
B, nh, L = 1, 37, 8192; h = nh * 128
qkv = torch.randn((B * L, 3 * h), device="cuda").bfloat16()
q, k, v = qkv[:, :h], qkv[:, h:2*h], qkv[:, 2*h:]
d_o = torch.randn((B * L, h), device="cuda").bfloat16()
out, lse = ops.attn_fwd(q, k, v, B, nh, L, L)
dq = torch.zeros((B * L, h), device="cuda", dtype=torch.float32)
dk = torch.zeros((B * L, h), device="cuda").bfloat16(); dv = torch.zeros_like(dk)
lib.lib().vds_debug_attn_pair_mode(1)
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
NCTA = 2 * B * nh * ((L + 255) // 256)
tr = torch.zeros((128 + 4 * NCTA,), device="cuda", dtype=torch.int64)
lib.lib().vds_debug_attn_bwd_trace(tr.data_ptr())
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
torch.cuda.synchronize()
lib.lib().vds_debug_attn_bwd_trace(None)
tl = tr.cpu()[128:].view(NCTA, 4)
tr = tr[:128].view(2, 4, 16)
# per-SM timeline: gaps between one CTA's exit and the next CTA's entry / set-up on the same SM
by_sm = {}
for c in range(NCTA):
    by_sm.setdefault(int(tl[c, 0]), []).append((int(tl[c, 1]), int(tl[c, 2]), int(tl[c, 3])))
t_first = min(int(x) for x in tl[:, 1]); t_last = max(int(x) for x in tl[:, 3])
gaps, setups, lives = [], [], []
for sm, lst in by_sm.items():
    lst.sort()
    for a, b in zip(lst[:-1], lst[1:]):
        gaps.append(b[0] - a[2])
    for e in lst:
        setups.append(e[1] - e[0]); lives.append(e[2] - e[0])
import statistics as S
print(f"kernel span {1e-3 * (t_last - t_first):.1f} us over {len(by_sm)} SMs, {NCTA // len(by_sm)} CTAs / SM; CTA life mean {1e-3 * S.mean(lives):.1f} us "
      f"(min {1e-3 * min(lives):.1f}, max {1e-3 * max(lives):.1f}); set-up mean {1e-3 * S.mean(setups):.2f} us (first-wave / later differ: min {1e-3 * min(setups):.2f} max {1e-3 * max(setups):.2f}); "
      f"exit -> next entry gap mean {1e-3 * S.mean(gaps):.2f} us (min {1e-3 * min(gaps):.2f}, max {1e-3 * max(gaps):.2f})")
st = tr.cpu()[:, 3]
snames = ["setup done", "K/V landed", "last dV/dK issued", "last dQ issued", "mma_done seen", "dK epilogue done",
          "last dq reduce issued", "dq reduces complete", "compute loop left", "after __syncthreads", "exit"]
for c in (0, 1):
    print(f"CTA {c} phase stamps (cycles since kernel entry): " + ", ".join(f"{n} {int(st[c, i])}" for i, n in enumerate(snames)))
t = tr.cpu().float() / 128.0     # per sub-tile
names = {0: ["w s_ready", "", "w dp_read", "w dvdk_ready", "", "w dq_ready", "", "", "S issue(+w)", "dP issue(+w)",
             "dQ issue(+w)", "dVdK issue(+w)", "", "", "", "TOTAL"],
         1: ["w s_full", "w dp_full", "w dq_full(i-2)", "", "exp phase", "", "tail total", "tail: wait+stores+stats", "tail: tmem_st_wait", "tail: proxy fence", "tmem_ld S (2 x32)", "tmem_ld dP (2 x32, sum)", "", "", "", "TOTAL"],
         2: ["w dq_full", "", "", "", "", "", "", "", "", "", "", "", "", "", "", "TOTAL"]}
for c in (0, 1):
    for role, rn in ((0, "issuer"), (1, "compute warp 4"), (2, "drain warp 8")):
        row = t[c, role]
        if row[15] == 0:
            continue
        print(f"CTA {c} {rn}: " + ", ".join(f"{n} {row[i]:.0f}" for i, n in enumerate(names[role]) if n))
