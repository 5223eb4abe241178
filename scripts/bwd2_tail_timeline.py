"""Timeline of the SPLIT launch of the CTA-pair attention backward at the bench shape (library built with -DVDS_B2_PROF,
run with VDS_B2_PROF_TAIL=1): which SM ran which piece when."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vds_b200
from vds_b200 import ops, lib
B, nh, L = 2, 4, 8208; h = nh * 128
qkv = torch.randn((B * L, 3 * h), device="cuda").bfloat16()
q, k, v = qkv[:, :h], qkv[:, h:2*h], qkv[:, 2*h:]
d_o = torch.randn((B * L, h), device="cuda").bfloat16()
out, lse = ops.attn_fwd(q, k, v, B, nh, L, L)
dq = torch.zeros((B * L, h), device="cuda", dtype=torch.float32)
dk = torch.zeros((B * L, h), device="cuda").bfloat16(); dv = torch.zeros_like(dk)
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
NCTA = 1024
tr = torch.zeros((128 + 4 * NCTA,), device="cuda", dtype=torch.int64)
lib.lib().vds_debug_attn_bwd_trace(tr.data_ptr())
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
torch.cuda.synchronize()
lib.lib().vds_debug_attn_bwd_trace(None)
buf = (ctypes.c_uint32 * 256)()
uni = int(os.environ.get("VDS_BWD2_UNIFORM", "0"))
n = lib.lib().vds_attn_bwd_tail_plan(42, 129, 74, buf, 256)
lens = [buf[i] >> 21 for i in range(n)] if uni == 0 else [43] * (42 * uni)
tl = tr.cpu()[128:].view(NCTA, 4)
rows = [(int(tl[c, 0]), int(tl[c, 1]), int(tl[c, 2]), int(tl[c, 3]), c) for c in range(NCTA) if int(tl[c, 1]) > 0]
t0 = min(r[1] for r in rows)
print(f"{len(rows)} CTAs, span {1e-3 * (max(r[3] for r in rows) - t0):.1f} us")
by_sm = {}
for sm, a, s_, e, c in rows:
    by_sm.setdefault(sm, []).append((a - t0, e - t0, c))
for sm in sorted(by_sm):
    if sm % 2: continue
    print(f"SM {sm:3d}: " + "  ".join(f"[cl {c // 2:3d} len {lens[c // 2] if c // 2 < len(lens) else -1:3d}: {1e-3 * a:6.1f} -> {1e-3 * e:6.1f}]" for a, e, c in sorted(by_sm[sm])))
