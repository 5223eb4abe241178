"""Own parameter sharding that replaces the reference's FSDP2 wrap (``apply_fsdp``, /root/reference/model.py:512-542).

Same grouping as the reference (one group per DiTBlock + one root group, model.py:523-541), same mixed
precision (bf16 parameter all-gather, fp32 gradient reduce-scatter averaged over ranks, train.py:323-325),
but laid out for one flat buffer per kind so that every collective is ONE NCCL call with no copy-in/out:

  master  fp32 [sum_g S_g]      this rank's shard of every group (+ Adam m, v alongside, see optim.py)
  shard16 bf16 [sum_g S_g]      bf16 image of the master shard (written by the fused AdamW kernel)
  full16  bf16 [sum_g W*S_g]    gathered compute parameters; kernels read shaped views of this
  gfull   fp32 [sum_g W*S_g]    gradient accumulation target of the wgrad GEMM epilogues
  gshard  fp32 [sum_g S_g]      reduce-scattered (averaged) gradients  (== gfull when W == 1)

Parameters stay gathered from one optimizer step to the next (result-identical to FSDP2's
reshard/re-gather, half the all-gather traffic; 180 GB of HBM makes this free).  All-gathers run on a
side stream right after the optimizer step, block 0 first, and the forward waits per group; the
reduce-scatter of block i runs on the side stream while block i-1 is still in backward.
"""
import os

import torch
import torch.distributed as dist

from . import ops

ALIGN = 8  # elements: keeps every bf16 parameter view 16-byte aligned (TMA requirement)


def _round_up(x, m):
    return (x + m - 1) // m * m


class Layout:
    """Pure-python layout math (device-free; unit-tested on CPU).

    Groups = one per DiTBlock + one root group (the reference's FSDP grouping, model.py:523-541).  With ``ckv_group = G``
    > 0 the ``context_kv`` Linear of every block is moved out of its block into extra groups of G consecutive blocks,
    laid out back to back ([weights of the G blocks][biases]): the K = 4096 GEMM of ``model.py:149-155`` has the same A
    operand (the caption embedding) in every block, so G blocks' worth run as ONE ``[B*Lc, 4096] x [4096, G*2h]`` GEMM
    (and one wgrad) on a contiguous ``[G*2h, 4096]`` weight.  Results are identical."""

    def __init__(self, named_shapes, depth, world, ckv_group=0):
        self.world = world
        self.depth = depth
        self.ckv_group = int(ckv_group) if ckv_group and ckv_group > 0 else 0
        self.n_ckv = (depth + self.ckv_group - 1) // self.ckv_group if self.ckv_group else 0
        groups = [[] for _ in range(depth + 1 + self.n_ckv)]  # depth blocks + root + ckv groups
        ckv_w = [[] for _ in range(self.n_ckv)]
        ckv_b = [[] for _ in range(self.n_ckv)]
        for n, shape in named_shapes:
            if n.startswith("blocks."):
                bi = int(n.split(".")[1])
                if self.ckv_group and ".context_kv." in n:
                    (ckv_w if n.endswith(".weight") else ckv_b)[bi // self.ckv_group].append((bi, n, tuple(shape)))
                else:
                    groups[bi].append((n, tuple(shape)))
            else:
                groups[depth].append((n, tuple(shape)))
        for gi in range(self.n_ckv):
            groups[depth + 1 + gi] = [(n, sh) for _, n, sh in sorted(ckv_w[gi])] + [(n, sh) for _, n, sh in sorted(ckv_b[gi])]
        self.groups = groups
        self.param = {}  # name -> (group, offset_in_group, numel, shape)
        self.group_numel = []  # padded, multiple of ALIGN*world
        for g, plist in enumerate(groups):
            off = 0
            for n, shape in plist:
                numel = 1
                for s in shape:
                    numel *= s
                self.param[n] = (g, off, numel, shape)
                off = _round_up(off + numel, ALIGN)
            self.group_numel.append(_round_up(max(off, ALIGN), ALIGN * world))
        self.shard_numel = [n // world for n in self.group_numel]
        self.full_base = [0]
        self.shard_base = [0]
        for g in range(len(groups)):
            self.full_base.append(self.full_base[-1] + self.group_numel[g])
            self.shard_base.append(self.shard_base[-1] + self.shard_numel[g])
        self.full_total = self.full_base[-1]
        self.shard_total = self.shard_base[-1]

    @property
    def n_groups(self):
        return len(self.groups)

    def ckv_of_block(self, i):
        """(ckv group index gi, position j of block i inside it, blocks in the group) or None."""
        if not self.ckv_group:
            return None
        gi = i // self.ckv_group
        nb = min(self.ckv_group, self.depth - gi * self.ckv_group)
        return gi, i - gi * self.ckv_group, nb

    def ckv_ranges(self, gi):
        """Full-buffer (start, rows, cols) of the stacked [nb*2h, Dc] weight of ckv group gi and (start, numel) of its
        stacked bias (None when the blocks have no context_kv bias).  Contiguity is asserted."""
        plist = self.groups[self.depth + 1 + gi]
        ws = [(n, sh) for n, sh in plist if n.endswith(".weight")]
        bs = [(n, sh) for n, sh in plist if n.endswith(".bias")]
        g = self.depth + 1 + gi
        start = self.full_base[g] + self.param[ws[0][0]][1]
        rows, cols = ws[0][1]
        pos = start
        for n, sh in ws:
            assert self.full_base[g] + self.param[n][1] == pos and sh == (rows, cols), "context_kv weights not contiguous"
            pos += rows * cols
        bias = None
        if bs:
            b0 = self.full_base[g] + self.param[bs[0][0]][1]
            pos = b0
            for n, sh in bs:
                assert self.full_base[g] + self.param[n][1] == pos, "context_kv biases not contiguous"
                pos += sh[0]
            bias = (b0, pos - b0)
        return (start, rows * len(ws), cols), bias

    def full_range(self, name):
        g, off, numel, _ = self.param[name]
        return self.full_base[g] + off, numel

    def shard_range(self, name, rank):
        """(start in this rank's shard space, start inside the parameter, length) of the part rank owns."""
        g, off, numel, _ = self.param[name]
        S = self.shard_numel[g]
        lo, hi = max(off, rank * S), min(off + numel, (rank + 1) * S)
        if hi <= lo:
            return self.shard_base[g], 0, 0
        return self.shard_base[g] + lo - rank * S, lo - off, hi - lo

    def adam_chunks(self, rank, group_of, chunk=4096):
        """Chunk table over this rank's shard space: lists (start, len, group id); chunks never straddle tensors."""
        starts, lens, gids = [], [], []
        for n in self.param:
            s, _, ln = self.shard_range(n, rank)
            gid = group_of.get(n)
            if gid is None:
                continue
            for c in range(0, ln, chunk):
                starts.append(s + c)
                lens.append(min(chunk, ln - c))
                gids.append(gid)
        return starts, lens, gids


class FlatShards:
    def __init__(self, model, param_dtype=torch.bfloat16, reduce_dtype=torch.float32, process_group=None,
                 device=None, ckv_group=None):
        assert param_dtype == torch.bfloat16 and reduce_dtype == torch.float32, \
            "the CUDA path computes in bf16 and reduces gradients in fp32 (train.py:323-325)"
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.backend = dist.get_backend(process_group) if dist.is_initialized() else None
        named = [(n, p) for n, p in model.named_parameters()]
        if device is None:
            device = named[0][1].device
        self.device = torch.device(device)
        self.depth = model.depth
        if ckv_group is None:
            # one grouped GEMM for the whole model at world size 1; groups of 4 blocks when sharded, so their gradient
            # reduce-scatters still overlap the backward of the blocks that follow (env VDS_CKV_GROUP overrides; 0 = off)
            ckv_group = int(os.environ.get("VDS_CKV_GROUP", model.depth if self.world == 1 else 4))
        has_ckv = any(".context_kv." in n for n, _ in named)
        self.layout = Layout([(n, p.shape) for n, p in named], model.depth, self.world,
                             ckv_group=ckv_group if has_ckv else 0)
        lay = self.layout
        f32 = dict(device=self.device, dtype=torch.float32)
        b16 = dict(device=self.device, dtype=torch.bfloat16)
        self.master = torch.zeros(lay.shard_total, **f32)
        self.full16 = torch.zeros(lay.full_total, **b16)
        # world size 1: the bf16 shard IS the gathered buffer (AdamW writes the compute copy in place, no gather copy)
        self.shard16 = self.full16 if self.world == 1 else torch.zeros(lay.shard_total, **b16)
        self.gfull = torch.zeros(lay.full_total, **f32)
        self.gshard = self.gfull if self.world == 1 else torch.zeros(lay.shard_total, **f32)
        # move the module's values into the master shard; re-point the module's Parameters at it
        self.params = {}
        for n, p in named:
            s, po, ln = lay.shard_range(n, self.rank)
            src = p.detach().to(device=self.device, dtype=torch.float32).reshape(-1)
            self.master[s:s + ln].copy_(src[po:po + ln])
            view = self.master[s:s + ln]
            if self.world == 1:
                view = view.view(p.shape)
            newp = torch.nn.Parameter(view, requires_grad=p.requires_grad)
            self._set_param(model, n, newp)
            self.params[n] = newp
        model._full_shapes = {n: lay.param[n][3] for n in lay.param}
        self._pview = {}
        self.grad_views = {}
        for n in lay.param:
            b, numel = lay.full_range(n)
            shape = lay.param[n][3]
            self._pview[n] = self.full16[b:b + numel].view(shape)
            self.grad_views[n] = self.gfull[b:b + numel].view(shape)
        self._seen_version = -1
        self.use_cuda = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1) if self.use_cuda else None
        self.ag_events = [None] * self.layout.n_groups
        self.rs_events = []
        self._fresh = False
        self._gather_pending = False
        self.refresh_from_master()

    @staticmethod
    def _set_param(model, name, newp):
        mod = model
        parts = name.split(".")
        for a in parts[:-1]:
            mod = getattr(mod, a)
        mod._parameters[parts[-1]] = newp

    # ------------------------------------------------------------------------------ collectives (plumbing)
    def _all_gather(self, out, inp):
        if self.world == 1:
            if out.data_ptr() != inp.data_ptr():
                out.copy_(inp)
            return
        if self.backend == "gloo":  # CPU tests
            parts = [torch.empty_like(inp) for _ in range(self.world)]
            dist.all_gather(parts, inp, group=self.pg)
            out.copy_(torch.cat(parts))
        else:
            dist.all_gather_into_tensor(out, inp, group=self.pg)

    def _reduce_scatter_avg(self, out, inp):
        if self.world == 1:
            return
        if getattr(self, "_accumulate", False):     # a second backward without zero_grad: out += reduce_scatter(inp)
            tmp = torch.empty_like(out)
            self._accumulate = False
            try:
                self._reduce_scatter_avg(tmp, inp)
            finally:
                self._accumulate = True
            out.add_(tmp)
            return
        if self.backend == "gloo":
            tmp = inp.clone()
            dist.all_reduce(tmp, group=self.pg)
            S = out.numel()
            out.copy_(tmp[self.rank * S:(self.rank + 1) * S] / self.world)
        else:
            dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.AVG, group=self.pg)

    def _group_slices(self, g):
        lay = self.layout
        fb, sb = lay.full_base[g], lay.shard_base[g]
        return slice(fb, fb + lay.group_numel[g]), slice(sb, sb + lay.shard_numel[g])

    # ------------------------------------------------------------------------------ parameters
    def refresh_from_master(self):
        """master (fp32 shard) -> bf16 shard -> all-gather (used at start-up, and whenever an optimizer other
        than ours wrote the master parameters)."""
        if self.use_cuda:
            ops.cast_f32_bf16(self.master, out=self.shard16)
        else:
            self.shard16.copy_(self.master)
        self.gather_params()
        self._seen_version = self.master._version

    def gather_params(self):
        """All-gather every group's bf16 shard on the side stream (root group first, then block 0, 1, ...)."""
        self._gather_pending = False
        order = self.group_order()
        if self.world == 1:
            return  # shard16 aliases full16
        if not self.use_cuda:
            for g in order:
                fs, ss = self._group_slices(g)
                self._all_gather(self.full16[fs], self.shard16[ss])
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            for g in order:
                fs, ss = self._group_slices(g)
                self._all_gather(self.full16[fs], self.shard16[ss])
                ev = torch.cuda.Event()
                ev.record(self.comm_stream)
                self.ag_events[g] = ev

    def group_order(self):
        """Order in which the forward first touches the groups: root, then per block its ckv group (if it opens one) and
        the block itself."""
        lay = self.layout
        order = [self.depth]
        for i in range(self.depth):
            c = lay.ckv_of_block(i)
            if c is not None and c[1] == 0:
                order.append(self.depth + 1 + c[0])
            order.append(i)
        return order

    def ckv_views(self, gi):
        """(stacked bf16 weight [nb*2h, Dc], stacked bf16 bias or None, fp32 grad of the weight, fp32 grad of the bias)
        of ckv group gi, as views of the gathered / gradient buffers."""
        (ws, rows, cols), bias = self.layout.ckv_ranges(gi)
        W = self.full16[ws:ws + rows * cols].view(rows, cols)
        gW = self.gfull[ws:ws + rows * cols].view(rows, cols)
        if bias is None:
            return W, None, gW, None
        b0, bn = bias
        return W, self.full16[b0:b0 + bn], gW, self.gfull[b0:b0 + bn]

    def mark_dirty(self):
        self._seen_version = -1

    def compute_params(self):
        if self.master._version != self._seen_version:
            self.refresh_from_master()
        elif self._gather_pending:
            # the last optimizer step left the gather to "the next step" (FusedAdamW.step(gather=False), used by the
            # graphed multi-GPU step): an eager forward that follows must not read the stale gathered copy
            self.gather_params()
        return self._pview

    def wait_group(self, g):
        """Called by the engine before it first touches group g's parameters (g == depth: root)."""
        ev = self.ag_events[g]
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self.ag_events[g] = None

    # ------------------------------------------------------------------------------ gradients
    def begin_backward(self):
        # "accumulate" = some trainable parameter still carries the .grad view a previous backward handed out (i.e. no
        # zero_grad in between).  Frozen parameters never get a .grad here (end_backward), so one frozen / optimizer-less
        # tensor cannot pin the buffer in accumulate mode.
        self._accumulate = any(p.grad is not None for p in self.params.values() if p.requires_grad)
        # W == 1: gfull IS the gradient (no reduce), accumulating = not zeroing it.  W > 1: gfull only holds this backward's
        # local gradient; the reduced shard gshard (what .grad views) adds each backward's reduce-scatter (_reduce_group).
        if not self._accumulate or self.world > 1:
            self.gfull.zero_()
        self.rs_events = []

    def _reduce_group(self, g):
        if self.world == 1:
            return
        fs, ss = self._group_slices(g)
        if not self.use_cuda:
            self._reduce_scatter_avg(self.gshard[ss], self.gfull[fs])
            return
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(done)
            self._reduce_scatter_avg(self.gshard[ss], self.gfull[fs])
            ev = torch.cuda.Event()
            ev.record(self.comm_stream)
            self.rs_events.append(ev)

    def block_backward_done(self, i):
        self._reduce_group(i)

    def ckv_backward_done(self, gi):
        self._reduce_group(self.depth + 1 + gi)

    def end_backward(self):
        self._reduce_group(self.depth)
        if self.use_cuda:
            for ev in self.rs_events:
                torch.cuda.current_stream().wait_event(ev)
        self.rs_events = []
        lay = self.layout
        for n, p in self.params.items():
            if n.endswith("blocks.0.lambda_param"):
                continue  # never used by block 0 (model.py:129): the reference leaves .grad = None too
            if not p.requires_grad:
                continue  # frozen: autograd would leave .grad = None as well
            if p.grad is None:
                s, _, ln = lay.shard_range(n, self.rank)
                p.grad = self.gshard[s:s + ln].view(p.shape)

    # ------------------------------------------------------------------------------ state_dict
    def full_state_dict(self):
        """Full, reference-shaped fp32 tensors (all-gathers the master shards when W > 1)."""
        lay = self.layout
        out = {}
        if self.world == 1:
            full = None
        for g in range(lay.n_groups):
            fs, ss = self._group_slices(g)
            buf = torch.empty(lay.group_numel[g], device=self.device, dtype=torch.float32)
            self._all_gather(buf, self.master[ss].contiguous())
            for n, _ in lay.groups[g]:
                _, off, numel, shape = lay.param[n]
                out[n] = buf[off:off + numel].view(shape).clone()
        return out


def get_device_mesh():
    """model.py:475-498 builds a (dp_replicate=1, dp_shard=world) mesh; here it is just the world group."""
    assert dist.is_initialized()
    return dist.group.WORLD


def apply_fsdp(dit_model, param_dtype, reduce_dtype):
    """Drop-in for model.py:512-542: shards every parameter group over all ranks (or just flattens them at
    world size 1, where the reference's own apply_fsdp raises NameError, model.py:489) and returns the model."""
    flat = FlatShards(dit_model, param_dtype, reduce_dtype)
    dit_model._flat = flat
    # a stock optimizer (e.g. the reference's optim.AdamW(fused=True), train.py:340-344) writes the fp32 master
    # views directly: refresh the bf16 compute copy before the next forward.  Our FusedAdamW does it itself.
    from torch.optim.optimizer import register_optimizer_step_post_hook
    from .optim import FusedAdamW
    owned = {id(p) for p in flat.params.values()}

    def _post_step(opt, args, kwargs):
        if isinstance(opt, FusedAdamW):
            return
        if any(id(p) in owned for g in opt.param_groups for p in g["params"]):
            flat.mark_dirty()
    dit_model._post_step_hook = register_optimizer_step_post_hook(_post_step)
    if flat.world > 1:
        def _hook(module, state_dict, prefix, local_metadata):
            if module is dit_model:
                for n, t in flat.full_state_dict().items():
                    state_dict[prefix + n] = t
            return state_dict
        dit_model._register_state_dict_hook(_hook)
    return dit_model
