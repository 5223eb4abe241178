"""apply_fsdp replacement (flat shards) + fused AdamW at world size 1 on the GPU."""
import pytest
import torch

from helpers import cos_sim, golden_case, params_of
from oracle import dit_oracle as O

pytestmark = pytest.mark.gpu
CONST = ["patch_proj", "context_kv", "positional_embedding"]


def test_flat_step_and_fused_adamw(cuda_dev):
    from vds_b200 import train
    from vds_b200.model import apply_fsdp
    from vds_b200.optim import FusedAdamW
    fx, cfg, model, (latent, noise, context, t) = golden_case("tiny_bias")
    model = model.to(cuda_dev)
    before = params_of(model)
    names = [n for n, _ in model.named_parameters()]
    model = apply_fsdp(model, torch.bfloat16, torch.float32)
    assert [n for n, _ in model.named_parameters()] == names
    sd = model.state_dict()
    assert all(torch.equal(sd[n], before[n]) for n in names)
    groups, settings = model.get_mup_setup(2 ** -7, 1e-1, CONST)
    assert len(groups) == fx["n_groups"]
    opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)
    latent, noise, context, t = [a.to(cuda_dev) for a in (latent, noise, context, t)]

    ref_state = {n: (before[n].clone(), torch.zeros_like(before[n]), torch.zeros_like(before[n])) for n in names}
    for step in (1, 2):
        opt.zero_grad()
        torch.manual_seed(fx["seeds"]["rope"])
        loss, _ = train.forward(model, latent, context, t=t, noise=noise, caption_dropout=0.0)
        loss.backward()
        grads = {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in model.named_parameters()}
        if step == 1:
            # gradients of the flat path agree with the golden reference run
            assert abs(loss.item() - fx["loss"]) <= 1e-2 * abs(fx["loss"])
            for n, g in fx["grads"].items():
                if g is not None:
                    assert cos_sim(grads[n].flatten().cpu()[g["idx"]], g["val"]) >= 0.995, n
        opt.step()
        torch.cuda.synchronize()
        # the AdamW kernel against the oracle's AdamW restatement on the SAME gradients
        for n in names:
            if grads[n] is None:
                continue
            lr, wd = settings[n]["lr"], settings[n]["wd"]
            p, m, v = ref_state[n]
            ref_state[n] = O.adamw_step(p, grads[n], m, v, step, lr, wd)
            got = dict(model.named_parameters())[n].detach()
            assert torch.allclose(got, ref_state[n][0], rtol=2e-5, atol=1e-7), (n, step)
    # bf16 compute copy was refreshed by the optimizer kernel
    P = model._flat.compute_params()
    for n in names:
        assert torch.equal(P[n], dict(model.named_parameters())[n].detach().bfloat16()), n


def test_stock_torch_adamw_on_flat_params(cuda_dev):
    """The reference's own optimizer construction (train.py:340-344) works on the flat parameters."""
    from vds_b200 import train
    from vds_b200.model import apply_fsdp
    fx, cfg, model, (latent, noise, context, t) = golden_case("tiny_nobias")
    model = apply_fsdp(model.to(cuda_dev), torch.bfloat16, torch.float32)
    groups, _ = model.get_mup_setup(2 ** -7, 1e-1, CONST)
    opt = torch.optim.AdamW(groups, betas=(0.95, 0.99), fused=True)
    latent, noise, context, t = [a.to(cuda_dev) for a in (latent, noise, context, t)]
    losses = []
    for step in range(3):
        opt.zero_grad()
        torch.manual_seed(fx["seeds"]["rope"])
        loss, _ = train.forward(model, latent, context, t=t, noise=noise, caption_dropout=0.0)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[2] < losses[0], losses   # it trains, i.e. the bf16 compute copy follows the master weights


def test_graphed_train_step_matches_eager(cuda_dev):
    """CUDA-graph captured step (device-side RoPE offsets + optimizer scalars) == eager step, step for step."""
    from vds_b200 import train
    from vds_b200.model import apply_fsdp
    from vds_b200.optim import FusedAdamW
    results = []
    for graphed in (False, True):
        fx, cfg, model, (latent, noise, context, t) = golden_case("tiny_bias")
        model = apply_fsdp(model.to(cuda_dev), torch.bfloat16, torch.float32)
        groups, _ = model.get_mup_setup(2 ** -7, 1e-1, CONST)
        opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)
        latent, noise, context, t = [a.to(cuda_dev) for a in (latent, noise, context, t)]
        stepper = train.GraphedTrainStep(model, opt, latent.shape, context.shape, device=cuda_dev, warmup=1) if graphed else None
        losses = []
        for step in range(4):
            torch.manual_seed(100 + step)
            for g in opt.param_groups:          # a moving learning rate, as a scheduler would produce
                g["lr"] = g["lr"] * 0.9
            if graphed:
                loss = stepper(latent, context, t, noise, caption_dropout=0.0)
            else:
                opt.zero_grad()
                loss, _ = train.forward(model, latent, context, t=t, noise=noise, caption_dropout=0.0)
                loss.backward()
                opt.step()
            losses.append(loss.item())
        torch.cuda.synchronize()
        results.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}))
        init = {n: p.detach().clone() for n, p in golden_case("tiny_bias")[2].named_parameters()}
    (l0, p0), (l1, p1) = results
    assert l1[2] != l1[1]                         # the replayed steps really see new offsets / updated weights
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-3 * abs(a), (l0, l1)
    # Adam turns near-zero gradients into +-lr steps, and the fp32 atomics of split-K / column sums are unordered,
    # so two runs agree on the update direction, not bit for bit: compare the accumulated updates.
    for n in p0:
        u0, u1 = (p0[n] - init[n].to(cuda_dev)).flatten(), (p1[n] - init[n].to(cuda_dev)).flatten()
        if u0.abs().max().item() == 0:
            assert u1.abs().max().item() == 0, n
            continue
        assert cos_sim(u0, u1) > 0.98, (n, cos_sim(u0, u1))


def test_graphed_step_stage_replay_pipeline(cuda_dev):
    """`stage()` issued behind a running `replay()` (what bench.py's e2e leg and a pipelined training loop do) gives the
    same losses as the one-call form: the staged copies queue behind the replay that still reads the old inputs, and the
    static loss tensor of step i is read back before step i+1 is launched."""
    from vds_b200 import train
    from vds_b200.model import apply_fsdp
    from vds_b200.optim import FusedAdamW
    results = []
    for pipelined in (False, True):
        fx, cfg, model, (latent, noise, context, t) = golden_case("tiny_bias")
        model = apply_fsdp(model.to(cuda_dev), torch.bfloat16, torch.float32)
        groups, _ = model.get_mup_setup(2 ** -7, 1e-1, CONST)
        opt = FusedAdamW(groups, betas=(0.95, 0.99), flat=model._flat)
        latent, noise, context, t = [a.to(cuda_dev) for a in (latent, noise, context, t)]
        batches = [(latent * (1.0 + 0.1 * i), context * (1.0 - 0.05 * i)) for i in range(6)]   # every step sees new inputs
        stepper = train.GraphedTrainStep(model, opt, latent.shape, context.shape, device=cuda_dev, warmup=1)
        losses = []
        if not pipelined:
            for i, (la, cx) in enumerate(batches):
                torch.manual_seed(200 + i)
                losses.append(stepper(la, cx, t, noise, caption_dropout=0.0).item())
        else:
            for i in range(2):                      # warm-up + capture through the one-call form
                torch.manual_seed(200 + i)
                losses.append(stepper(batches[i][0], batches[i][1], t, noise, caption_dropout=0.0).item())
            assert stepper.graph is not None
            torch.manual_seed(202)
            stepper.stage(batches[2][0], batches[2][1], t, noise, caption_dropout=0.0)
            for i in range(2, 6):
                loss = stepper.replay()
                if i + 1 < 6:                       # host side of the next step, queued behind the running replay
                    torch.manual_seed(200 + i + 1)
                    stepper.stage(batches[i + 1][0], batches[i + 1][1], t, noise, caption_dropout=0.0)
                    # early read-back over the side stream (the backward graph of step i may still be running): it must
                    # see step i's loss although step i+1's inputs are already queued
                    losses.append(stepper.loss_value())
                    assert abs(losses[-1] - loss.item()) == 0.0
                else:
                    losses.append(stepper.loss_value())
        results.append(losses)
    l0, l1 = results
    assert len(l0) == len(l1) == 6
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 2e-3 * abs(a), (l0, l1)
    assert len({round(x, 4) for x in l0}) > 3       # the steps really differ
