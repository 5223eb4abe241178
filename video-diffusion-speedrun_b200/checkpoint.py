"""Checkpoint I/O in the reference's layout.

The reference saves ``get_model_state_dict(dit_model)`` with ``torch.distributed.checkpoint.save`` into
``checkpoints/<run>/<step>`` (``train.py:553-584``) and loads by converting that directory with
``dcp_to_torch_save`` into ``temp.pt`` followed by ``load_state_dict(assign=True)`` — before ``apply_fsdp`` when
training resumes (``train.py:293-320``), on a meta-constructed bf16 model when sampling (``sample.py:32-65``).
The functions below write and read exactly that format, so checkpoints move between the reference and this
implementation in both directions (state_dict keys and shapes are the reference's).

``skip_rope=True`` leaves the two persistent RoPE tables (``rope.freqs_hwt_{cos,sin}``, 2 x 128^3 x hd/2 fp32 = 1 GiB at
hd = 128) out of the file; ``load_checkpoint`` regenerates them from ``ThreeDimRotary``'s constructor (they are a pure
function of the head dim), so ``strict=True`` loads keep working.
"""
import os
from pathlib import Path

import torch
import torch.distributed as dist

ROPE_KEYS = ("rope.freqs_hwt_cos", "rope.freqs_hwt_sin")


def _is_master():
    return (not dist.is_initialized()) or dist.get_rank() == 0


def save_checkpoint(model, path, skip_rope=False):
    """All ranks call this (the state_dict of a sharded model is assembled collectively; DCP de-duplicates the
    replicated tensors).  Returns the list of saved keys."""
    import torch.distributed.checkpoint as dcp
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    if skip_rope:
        for k in ROPE_KEYS:
            sd.pop(k, None)
    path = Path(path)
    if _is_master():
        os.makedirs(path, exist_ok=True)
    if dist.is_initialized():
        dist.barrier()
    dcp.save(sd, checkpoint_id=path)
    return sorted(sd)


def read_checkpoint(path, map_location="cpu"):
    """DCP directory -> plain state_dict through ``temp.pt`` like the reference (train.py:299-304, sample.py:34-39)."""
    from torch.distributed.checkpoint.format_utils import dcp_to_torch_save
    path = Path(path)
    temp = path / "temp.pt"
    if _is_master() and not temp.exists():
        dcp_to_torch_save(path, temp)
    if dist.is_initialized():
        dist.barrier()
    sd = torch.load(temp, map_location=map_location)
    return {k.replace("module.", "").replace("_orig_mod.", ""): v for k, v in sd.items()}


def load_checkpoint(model, path, device="cpu", dtype=torch.float32, strict=True):
    """``load_state_dict(assign=True)`` of a checkpoint directory into ``model`` (which may live on the meta device,
    sample.py:41-61).  Call before ``apply_fsdp`` (train.py:293-325).  Returns the load status."""
    sd = {k: v.clone().to(device, dtype=dtype if v.is_floating_point() else None) for k, v in read_checkpoint(path).items()}
    rope = getattr(model, "rope", None)
    if rope is not None and any(k not in sd for k in ROPE_KEYS):
        from .model import ThreeDimRotary
        fresh = ThreeDimRotary(model.hidden_size // (2 * model.num_heads), h=128, w=128, t=128)   # model.py:310-314
        for k in ROPE_KEYS:
            if k not in sd:
                sd[k] = getattr(fresh, k.split(".", 1)[1]).clone().to(device, dtype=dtype)
    status = model.load_state_dict(sd, assign=True, strict=strict)
    return status
