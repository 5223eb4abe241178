"""torchrun worker: sharded (W ranks, NCCL) train steps must equal the single-GPU flat path on the averaged
gradients.  Launched by tests/test_multi_gpu.py (or by hand: torchrun --nproc-per-node 2 tests/mgpu_worker.py)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vds_b200  # noqa: E402,F401
from helpers import build_model, cos_sim  # noqa: E402
from oracle import dit_oracle as O  # noqa: E402
from vds_b200 import train  # noqa: E402
from vds_b200.model import apply_fsdp  # noqa: E402
from vds_b200.optim import FusedAdamW  # noqa: E402

CONST = ["patch_proj", "context_kv", "positional_embedding"]
# VDS_MGPU_CFG=debug: the run_debug.sh width (512, 4 heads x 128, T5-width context) at depth 3 on a [4,16,16] latent;
# default: the tiny 256-wide model.  VDS_MGPU_GRAPH=1: the sharded steps go through train.GraphedTrainStep (NCCL
# all-gathers / reduce-scatters captured inside the CUDA graph) instead of the Python-issued step.
DEBUG_W = os.environ.get("VDS_MGPU_CFG", "tiny") == "debug"
GRAPH = os.environ.get("VDS_MGPU_GRAPH", "0") == "1"
NSTEPS = 4 if GRAPH else 2
if DEBUG_W:
    CFG = dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=512, depth=3, num_heads=4, mlp_ratio=4.0,
               cross_attn_input_size=4096, residual_v=True, train_bias_and_rms=False, use_rope=True)
    DATA = (2, (4, 16, 16), 512, 4096)
else:
    CFG = dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=256, depth=3, num_heads=2, mlp_ratio=4.0,
               cross_attn_input_size=64, residual_v=True, train_bias_and_rms=True, use_rope=True)
    DATA = (2, (4, 8, 8), 24, 64)


def data(rank, dev):
    return [a.to(dev) for a in O.make_inputs(CFG, DATA[0], DATA[1], DATA[2], DATA[3], 100 + rank)]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("VDS_MGPU_DEBUG") == "1":
        import faulthandler
        faulthandler.dump_traceback_later(75, exit=False)     # where is every rank if the run wedges
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model = apply_fsdp(build_model(CFG, 0, 1).to(dev), torch.bfloat16, torch.float32)
    groups, settings = model.get_mup_setup(2 ** -7, 1e-1, CONST)
    opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)
    latent, noise, context, t = data(rank, dev)
    losses = []
    stepper = train.GraphedTrainStep(model, opt, latent.shape, context.shape, device=dev, warmup=1) if GRAPH else None
    dbg = os.environ.get("VDS_MGPU_DEBUG") == "1"
    for step in range(NSTEPS):
        if dbg:
            torch.cuda.synchronize()
            print(f"[rank {rank}] step {step} begin", flush=True)
        torch.manual_seed(50 + step)
        if stepper is not None:     # step 0 eager on the capture stream, step 1 captures + replays, steps 2.. replay
            loss = stepper(latent, context, t, noise, caption_dropout=0.0).clone()
        else:
            opt.zero_grad()
            loss, _ = train.forward(model, latent, context, t=t, noise=noise, caption_dropout=0.0)
            loss.backward()
            opt.step()
        losses.append(loss.detach())
    torch.cuda.synchronize()
    if dbg:
        print(f"[rank {rank}] steps done", flush=True)
    if stepper is not None:
        assert stepper.graph is not None, "the multi-GPU step was not captured"
    sd = model.state_dict()  # full tensors (all-gathers the fp32 master shards)
    all_loss = [torch.zeros(NSTEPS, device=dev) for _ in range(world)]
    dist.all_gather(all_loss, torch.stack(losses))
    dist.barrier()
    ok = True
    if rank == 0:
        # single-process reference: same model, every rank's batch in turn, gradients averaged in fp32
        ref = build_model(CFG, 0, 1).to(dev)
        names = [n for n, _ in ref.named_parameters()]
        state = {n: (p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)) for n, p in ref.named_parameters()}
        for step in range(NSTEPS):
            gsum = {}
            for r in range(world):
                la, no, cx, tt = data(r, dev)
                ref.zero_grad(set_to_none=True)
                torch.manual_seed(50 + step)
                loss, _ = train.forward(ref, la, cx, t=tt, noise=no, caption_dropout=0.0)
                loss.backward()
                if abs(loss.item() - all_loss[r][step].item()) > 2e-3 * abs(loss.item()):
                    print(f"loss mismatch step {step} rank {r}: {loss.item()} vs {all_loss[r][step].item()}")
                    ok = False
                for n, p in ref.named_parameters():
                    if p.grad is not None:
                        gsum[n] = gsum.get(n, 0) + p.grad.float() / world
            with torch.no_grad():
                for n, p in ref.named_parameters():
                    if n not in gsum:
                        continue
                    lr, wd = settings[n]["lr"], settings[n]["wd"]
                    state[n] = O.adamw_step(state[n][0], gsum[n], state[n][1], state[n][2], step + 1, lr, wd)
                    p.copy_(state[n][0])
        worst = 1.0
        for n in names:
            c = cos_sim(sd[n], state[n][0])
            worst = min(worst, c)
            rel = (sd[n].float() - state[n][0]).abs().max().item() / (state[n][0].abs().max().item() + 1e-12)
            if c < 0.9999 or rel > 2e-2:
                print(f"param mismatch {n}: cos {c:.6f} rel {rel:.3e}")
                ok = False
        print(f"MGPU world={world} cfg={'debug' if DEBUG_W else 'tiny'} graph={int(GRAPH)}: worst param cosine after "
              f"{NSTEPS} sharded steps vs single-GPU reference {worst:.6f}; {'OK' if ok else 'FAIL'}")
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    if stepper is not None:
        stepper.close()      # a live CUDA graph holding NCCL kernels makes destroy_process_group wait forever
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
