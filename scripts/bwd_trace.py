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
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
tr = torch.zeros((160, 8), device="cuda", dtype=torch.int64)
lib.lib().vds_debug_attn_bwd_trace(tr.data_ptr())
ops.attn_bwd(q, k, v, out, d_o, lse, B, nh, L, L, dq, dk=dk, dv=dv)
torch.cuda.synchronize()
lib.lib().vds_debug_attn_bwd_trace(None)
t = tr.cpu()
t0 = t[0, 0].item()
names = ["sdp_issue", "mma_issue", "cmp_start", "math_done", "pds_arrive", "drain_start", "drained", "s_issued"]
print("iter " + " ".join(f"{n:>11s}" for n in names))
for i in list(range(0, 6)) + list(range(30, 36)):
    print(f"{i:4d} " + " ".join(f"{(t[i, s].item() - t0):11d}" for s in range(8)))
d = t[10:50]
print("mean period (cycles):", (d[-1, 1] - d[0, 1]).item() / 39)
print("compute: start->math_done", (d[:, 3] - d[:, 2]).float().mean().item(), " math_done->pds", (d[:, 4] - d[:, 3]).float().mean().item())
print("mma_issue(i) - pds_arrive(i)", (d[:, 1] - d[:, 4]).float().mean().item())
print("drain_start(i) - mma_issue(i)", (d[:, 5] - d[:, 1]).float().mean().item(), " drained - drain_start", (d[:, 6] - d[:, 5]).float().mean().item())
print("cmp_start(i) - sdp_issue(i)", (d[:, 2] - d[:, 0]).float().mean().item())
print("sdp_issue(i+1) - mma_issue(i)", (d[1:, 0] - d[:-1, 1]).float().mean().item())
print("cmp_start(i+1) - pds_arrive(i)", (d[1:, 2] - d[:-1, 4]).float().mean().item())
