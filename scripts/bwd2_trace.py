"""Per-sub-tile clock64 trace of cluster 0 of the CTA-pair attention backward (attention_bwd2.cu: B2_TRACE)."""
import sys, os
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
lib.lib().vds_debug_attn_pair_mode(1)
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
tr = torch.zeros((2, 160, 16), device="cuda", dtype=torch.int64)
lib.lib().vds_debug_attn_bwd_trace(tr.data_ptr())
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
torch.cuda.synchronize()
lib.lib().vds_debug_attn_bwd_trace(None)
t = tr.cpu()
names = ["S_wait", "rows_ok", "S_issue", "S_done_i", "dP_issue", "pds_wait", "pds_ok", "drn_ok", "dQ_issue", "c_seesS", "exp_done",
         "c_seesdP", "math_done", "dsfree", "pds_arr", "drained"]
for c in (0, 1):
    t0 = t[0, 0, 0].item()
    print(f"--- CTA {c} (cycles since leader's first stamp)")
    print("iter " + " ".join(f"{n:>9s}" for n in names))
    for i in list(range(0, 4)) + list(range(40, 46)):
        print(f"{i:4d} " + " ".join(f"{(t[c, i, s].item() - t0) if t[c, i, s].item() else 0:9d}" for s in range(16)))
d = t[0, 20:100]
f = lambda a: a.float().mean().item()
per = (d[-1, 6] - d[0, 6]).item() / (d.shape[0] - 1)
print("leader mean period (cycles):", per)
print("issuer S(k): rows wait", f(d[:, 1] - d[:, 0]), " stat wait", f(d[:, 2] - d[:, 1]))
print("issuer dP(k): dp_read wait", f(d[:, 4] - d[:, 3]))
print("issuer dQ(i-1): dP issue(i+1) -> dq wait start(i-1)", f(d[:-2, 7] - d[2:, 4]), " exchange+drain wait", f(d[:, 8] - d[:, 7]),
      " dQ issue(i-1) -> pds wait start(i)", f(d[1:, 5] - d[:-1, 8]))
print("issuer dV/dK(i): pds wait", f(d[:, 6] - d[:, 5]), " pds_ok(i) -> S wait start(i+2)", f(d[2:, 0] - d[:-2, 6]))
for c in (0, 1):
    e = t[c, 20:100]
    print(f"CTA {c} compute: seesS->exp_done", f(e[:, 10] - e[:, 9]), " wait dP", f(e[:, 11] - e[:, 10]), " dS math", f(e[:, 12] - e[:, 11]),
          " wait dq_full(i-2)", f(e[:, 13] - e[:, 12]), " stores+fence+arrive+copy", f(e[:, 14] - e[:, 13]), " pds_arr(i)->seesS(i+1)", f(e[1:, 9] - e[:-1, 14]))
    print(f"CTA {c}: pds_arr(i) -> leader pds_ok(i)", f(d[:, 6] - e[:, 14]), " leader dQ issue(i) -> drained(i)", f(e[:, 15] - d[:, 8]))
