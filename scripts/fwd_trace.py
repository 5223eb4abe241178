import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vds_b200
from vds_b200 import ops, lib
B, nh, L = 2, 4, 8208; h = nh * 128
qkv = torch.randn((B * L, 3 * h), device="cuda").bfloat16()
q, k, v = qkv[:, :h], qkv[:, h:2*h], qkv[:, 2*h:]
ops.attn_fwd(q, k, v, B, nh, L, L)
tr = torch.zeros((160, 8), device="cuda", dtype=torch.int64)
lib.lib().vds_debug_attn_bwd_trace(tr.data_ptr())
ops.attn_fwd(q, k, v, B, nh, L, L)
torch.cuda.synchronize()
lib.lib().vds_debug_attn_bwd_trace(None)
t = tr.cpu(); t0 = t[0, 2].item()
names = ["PV0_issue", "PV1_issue", "sm0_start", "sm0_max", "sm0_pub", "sm1_start", "sm1_max", "sm1_pub"]
print("iter " + " ".join(f"{n:>10s}" for n in names))
for i in list(range(0, 4)) + list(range(30, 36)):
    print(f"{i:4d} " + " ".join(f"{(t[i, s].item() - t0):10d}" for s in range(8)))
d = t[10:60]
print("period", (d[-1, 0] - d[0, 0]).item() / 49)
print("softmax0: start->max", (d[:, 3] - d[:, 2]).float().mean().item(), "max->publish", (d[:, 4] - d[:, 3]).float().mean().item())
print("softmax1: start->max", (d[:, 6] - d[:, 5]).float().mean().item(), "max->publish", (d[:, 7] - d[:, 6]).float().mean().item())
print("PV0_issue(j) - sm0_pub(j)", (d[:, 0] - d[:, 4]).float().mean().item(), " PV1_issue(j) - sm1_pub(j)", (d[:, 1] - d[:, 7]).float().mean().item())
print("sm0_start(j+1) - PV0_issue(j)", (d[1:, 2] - d[:-1, 0]).float().mean().item(), " sm1_start(j+1) - PV1_issue(j)", (d[1:, 5] - d[:-1, 1]).float().mean().item())
