"""Fused multi-tensor AdamW on the flat fp32 master shard — replaces ``optim.AdamW(groups, betas=(0.95, 0.99),
fused=True)`` of the reference (/root/reference/train.py:340-344, stepped at train.py:433).

One kernel launch updates every parameter of this rank's shard (per-group lr / weight-decay by value, so
the HF warm-up schedulers of train.py:349-362 keep working through ``param_groups[i]["lr"]``) and emits the
bf16 image that the next parameter all-gather sends.
"""
import ctypes

import torch

from . import lib as L


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, flat=None, fused=True):
        if flat is None:
            raise ValueError("FusedAdamW needs the FlatShards object of apply_fsdp(model) (model._flat)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.flat = flat
        if len(self.param_groups) > 16:
            raise ValueError("at most 16 parameter groups (the reference's muP setup has 6-7)")
        name_of = {id(p): n for n, p in flat.params.items()}
        group_of = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                group_of[name_of[id(p)]] = gi
        starts, lens, gids = flat.layout.adam_chunks(flat.rank, group_of)
        dev = flat.device
        self.n_chunks = len(starts)
        self.chunk_start = torch.tensor(starts, dtype=torch.int64, device=dev)
        self.chunk_len = torch.tensor(lens, dtype=torch.int32, device=dev)
        self.chunk_group = torch.tensor(gids, dtype=torch.int32, device=dev)
        self.m = torch.zeros_like(flat.master)
        self.v = torch.zeros_like(flat.master)
        self._step = 0
        self.hyper_dev = None   # set by train.GraphedTrainStep: device copy of [lr | wd | bc1 | sqrt(bc2)]

    @torch.no_grad()
    def step(self, closure=None, gather=True):
        """One fused AdamW update of this rank's shard.  `gather=False` skips the parameter all-gather that normally
        follows (train.GraphedTrainStep at world size > 1 issues it at the top of the next captured step instead)."""
        assert closure is None
        flat = self.flat
        self._step += 1
        n = len(self.param_groups)
        lr = (ctypes.c_float * n)(*[float(g["lr"]) for g in self.param_groups])
        wd = (ctypes.c_float * n)(*[float(g["weight_decay"]) for g in self.param_groups])
        b1, b2 = self.param_groups[0]["betas"]
        eps = self.param_groups[0]["eps"]
        L.check(L.lib().vds_adamw(flat.master.data_ptr(), flat.gshard.data_ptr(), self.m.data_ptr(),
                                  self.v.data_ptr(), flat.shard16.data_ptr(), self.chunk_start.data_ptr(),
                                  self.chunk_len.data_ptr(), self.chunk_group.data_ptr(), self.n_chunks, lr, wd, n,
                                  b1, b2, eps, self._step, 1.0,
                                  self.hyper_dev.data_ptr() if self.hyper_dev is not None else None,
                                  torch.cuda.current_stream().cuda_stream), "vds_adamw")
        if gather:
            flat.gather_params()  # side stream; the next forward waits per group
        else:
            flat._gather_pending = flat.world > 1
        return None

    def hyper_values(self, step):
        """[lr[16] | wd[16] | bc1 | sqrt(bc2)] for optimizer step `step` (1-based) from the current param_groups."""
        b1, b2 = self.param_groups[0]["betas"]
        lr = [float(g["lr"]) for g in self.param_groups] + [0.0] * (16 - len(self.param_groups))
        wd = [float(g["weight_decay"]) for g in self.param_groups] + [0.0] * (16 - len(self.param_groups))
        return lr + wd + [1.0 - b1 ** step, (1.0 - b2 ** step) ** 0.5]

    def zero_grad(self, set_to_none=True):
        # the flat gradient buffer is zeroed by the next backward; dropping .grad marks "fresh" (shard.py)
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None
