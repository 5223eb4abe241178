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
tr = torch.zeros((2, 3, 16), device="cuda", dtype=torch.int64)
lib.lib().vds_debug_attn_bwd_trace(tr.data_ptr())
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
torch.cuda.synchronize()
lib.lib().vds_debug_attn_bwd_trace(None)
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
