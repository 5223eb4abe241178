"""Mirror of the train-step arithmetic of the reference ``train.py`` (/root/reference/train.py:89-125,
429-435) on the fused CUDA path: timestep sampling, noise, z_t / v-target, model, MSE loss, backward.

``forward(dit_model, latent, caption_encoded, ...) -> (total_loss, diffusion_loss)`` keeps the reference's
name and return convention (train.py:51-145); the T5 encoder / dataloader (out of scope) are
replaced by the caller handing in ``caption_encoded`` directly.
"""
import torch

from . import engine, ops


def shift_time(t, alpha=8.0):
    """train.py:93-96."""
    return t * alpha / (1 + (alpha - 1) * t)


class TrainStepFunction(torch.autograd.Function):
    """loss = mse(v, DiT(z_t)) as ONE autograd node: z_t is formed inside the patch gather, the loss and
    its gradient come from one fused reduction kernel, the model backward is the hand-written engine."""

    @staticmethod
    def forward(ctx, model, need, latent, noise, context, t, starts_dev, *params):
        P = model._param_view()
        out, c = engine.forward(model, P, latent, context, t, save=need, noise=noise, rope_starts_dev=starts_dev)
        loss, _, lb = ops.loss_fwd_bwd(latent, noise, out, want_grad=False, want_batch=True)
        ctx.model, ctx.P, ctx.c = model, P, c
        ctx.saved = (latent, noise, out)
        ctx.n_params = len(params)
        ctx.dtypes = [p.dtype for p in params]
        ctx.mark_non_differentiable(lb)
        return loss.view(()), lb

    @staticmethod
    def backward(ctx, gloss, _glb):
        latent, noise, out = ctx.saved
        g = gloss.detach().to(torch.float32).reshape(1).contiguous()
        _, d_out, _ = ops.loss_fwd_bwd(latent, noise, out, want_grad=True, want_loss=False, grad_scale_dev=g)
        grads = engine.run_backward(ctx.model, ctx.P, ctx.c, d_out, ctx.dtypes)
        ctx.c = None
        return (None, None, None, None, None, None, None) + grads


CAPTION_DROPOUT = 0.01   # train.py:86


def drop_captions(caption_encoded, p=CAPTION_DROPOUT, out=None):
    """train.py:86-87: with probability 1 % the whole caption embedding of a sample is zeroed (this zero embedding is
    what sampling/sample.py:104 uses as the CFG negative).  The draw is ``torch.rand(B, device=device)`` on the global
    generator of the caption's device, like the reference.  Written as a mask fill (no ``nonzero`` host sync, so it can
    sit in front of a CUDA-graph replay); `out` = fill in place."""
    if p <= 0.0:
        return caption_encoded
    do_zero_out = torch.rand(caption_encoded.shape[0], device=caption_encoded.device) < p
    mask = do_zero_out.view(-1, *([1] * (caption_encoded.dim() - 1)))
    if out is not None:
        return out.masked_fill_(mask, 0)
    return caption_encoded.masked_fill(mask, 0)


def forward(dit_model, latent, caption_encoded, generator=None, t=None, noise=None, rope_starts_dev=None,
            caption_dropout=CAPTION_DROPOUT):
    """train.py:51-145 on the fused path.  latent [B,16,T,H,W], caption_encoded [B,512,4096] (bf16, CUDA).
    `t` / `noise` may be supplied (parity tests); otherwise drawn exactly like train.py:89-105.  The 1 % caption
    zero-out of train.py:86-87 is applied here (`caption_dropout=0` turns it off for parity tests)."""
    device = latent.device
    vae_latent = latent.to(torch.bfloat16).contiguous()  # train.py:73
    batch_size = vae_latent.size(0)
    caption_encoded = drop_captions(caption_encoded.to(torch.bfloat16), caption_dropout)  # train.py:84-87
    if t is None:
        z = torch.randn(batch_size, device=device, dtype=torch.bfloat16, generator=generator)
        t = shift_time(torch.sigmoid(z))
    if noise is None:
        noise = torch.randn(vae_latent.shape, device=device, dtype=torch.bfloat16, generator=generator)
    params = [p for _, p in dit_model.named_parameters()]
    need = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    loss, loss_batchwise = TrainStepFunction.apply(dit_model, need, vae_latent, noise.contiguous(),
                                                   caption_encoded.to(torch.bfloat16), t.to(torch.bfloat16), rope_starts_dev,
                                                   *params)
    dit_model.last_loss_batchwise = loss_batchwise
    return loss, loss


class GraphedTrainStep:
    """One whole train step (zero_grad -> fused forward/loss -> backward -> FusedAdamW) captured in a CUDA graph.

    Everything that changes from step to step lives in device buffers that are refreshed before each replay: the
    batch, the RoPE start offsets (still drawn from the global CPU generator in the reference's order h, w, t) and the
    optimizer scalars (per-group lr / wd, bias corrections).  ~1100 kernel launches become one graph launch, which is
    what the small workloads (DiT-B at 256x256, the S_small debug shape) are bound by.

    World size > 1: the graph also holds the collectives of shard.FlatShards on their side stream — the bf16 parameter
    all-gathers are issued at the TOP of the captured step (root group, block 0, 1, ...; the forward waits per group,
    so they overlap the forward exactly like the eager path's post-optimizer gathers overlap the next forward) and the
    per-block fp32 reduce-scatters are forked off the backward.  Every fork joins the capture stream again inside the
    step (the forward waits for every group, end_backward for every reduce-scatter), which is what stream capture
    requires.  The host issues ONE graph launch per step on every rank.
    """

    def __init__(self, model, optimizer, latent_shape, context_shape, device="cuda", warmup=2):
        from . import engine as _engine
        assert model._flat is not None, "GraphedTrainStep needs apply_fsdp(model) (flat parameter / gradient buffers)"
        self.model, self.opt, self._engine = model, optimizer, _engine
        dev = torch.device(device)
        bf = dict(device=dev, dtype=torch.bfloat16)
        self.latent = torch.zeros(latent_shape, **bf)
        self.noise = torch.zeros(latent_shape, **bf)
        self.context = torch.zeros(context_shape, **bf)
        self.t = torch.zeros((latent_shape[0],), **bf)
        self.starts_dev = torch.zeros(3, device=dev, dtype=torch.int32)
        self.hyper_dev = torch.zeros(34, device=dev, dtype=torch.float32)
        B, C, T, H, W = latent_shape
        self.thw = (T // model.time_patch_size, H // model.patch_size, W // model.patch_size)
        self.graph = None        # forward + loss
        self.graph_b = None      # backward + optimizer (same memory pool)
        self._carry = None       # tensors that live across the two graphs (d_out, saved activations, parameter views)
        self.loss = None
        self.warmup = warmup
        self.calls = 0
        self.launches_per_step = 0
        self._side = None
        self._loss_ready = None  # recorded between the two graph launches: the loss is final once the first graph is done
        self._rd_stream = None
        self._loss_host = None

    def close(self):
        """Drops the captured graph (and its private memory pool).  Required before ``dist.destroy_process_group()``
        at world size > 1: a live graph that contains NCCL kernels keeps the communicator busy and the destroy hangs."""
        self.graph = None
        self.graph_b = None
        self._carry = None
        self.loss = None
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _refresh_scalars(self):
        st, sh, sw = self._engine.draw_rope_starts(self.model.rope, self.thw)   # consumes the CPU RNG like the reference
        # Fresh pinned staging tensors every step: torch's caching host allocator does not hand a block out again
        # before the async copy that reads it has run, so the host may queue many steps ahead of the GPU.
        self.starts_dev.copy_(torch.tensor([st, sh, sw], dtype=torch.int32).pin_memory(), non_blocking=True)
        hyper = torch.tensor(self.opt.hyper_values(self.opt._step + 1), dtype=torch.float32).pin_memory()
        self.hyper_dev.copy_(hyper, non_blocking=True)

    def _fwd_body(self):
        """Forward + loss, without torch.autograd in the loop (the engine is called directly), so the capture contains only
        our kernels + memsets and no autograd-engine stream bookkeeping."""
        model, eng = self.model, self._engine
        sharded = model._flat.world > 1
        with torch.no_grad():
            self.opt.zero_grad()
            if sharded:
                model._flat.gather_params()      # side stream; consumed group by group by the forward below
            P = model._param_view()
            out, c = eng.forward(model, P, self.latent, self.context, self.t, save=True, noise=self.noise,
                                 rope_starts_dev=self.starts_dev)
            loss, d_out, lb = ops.loss_fwd_bwd(self.latent, self.noise, out, want_grad=True, want_batch=True)
            model.last_loss_batchwise = lb
        return loss.view(()), (P, c, d_out)

    def _bwd_body(self, carry):
        model, eng = self.model, self._engine
        P, c, d_out = carry
        with torch.no_grad():
            eng.run_backward(model, P, c, d_out, None)
            self.opt.step(gather=model._flat.world == 1)    # sharded: the next step's graph gathers at its top

    def _step_body(self):
        loss, carry = self._fwd_body()
        self._bwd_body(carry)
        return loss

    def stage(self, latent, context, t, noise, caption_dropout=CAPTION_DROPOUT):
        """Host side of one step: the inputs and the per-step scalars go into the graph's static buffers (stream-ordered
        copies).  May be issued while the previous replay is still running — it queues behind it — so a training loop can
        prepare step i+1 on the host while the GPU computes step i (``replay()`` launches what was staged last)."""
        self.latent.copy_(latent, non_blocking=True)
        self.context.copy_(context, non_blocking=True)
        drop_captions(self.context, caption_dropout, out=self.context)   # train.py:86-87, outside the graph
        self.t.copy_(t, non_blocking=True)
        self.noise.copy_(noise, non_blocking=True)
        self._refresh_scalars()

    def replay(self):
        """Runs the step on the staged inputs; returns the (static) loss tensor."""
        # The device copy of the optimizer scalars is installed only while this call captures / replays: a later eager
        # opt.step() (e.g. after falling back from the graph) must read lr / wd / bias corrections by value again.
        self.opt.hyper_dev = self.hyper_dev
        try:
            return self._run()
        finally:
            self.opt.hyper_dev = None

    def loss_value(self):
        """Host value of the last step's loss as soon as its forward graph has finished (the backward may still be
        running): the D2H read goes over a side stream that waits only for the event between the two graph launches."""
        if self.graph is None or self._loss_ready is None:
            return float(self.loss.item())          # warm-up steps (not captured yet)
        if self._rd_stream is None:
            self._rd_stream = torch.cuda.Stream()
            self._loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        self._rd_stream.wait_event(self._loss_ready)
        with torch.cuda.stream(self._rd_stream):
            self._loss_host.copy_(self.loss, non_blocking=True)
        self._rd_stream.synchronize()
        return float(self._loss_host)

    def __call__(self, latent, context, t, noise, caption_dropout=CAPTION_DROPOUT):
        self.stage(latent, context, t, noise, caption_dropout)
        return self.replay()

    def _run(self):
        if self.graph is None:
            # PyTorch's whole-network capture recipe: warm up on the side stream the capture will use (autograd's
            # stream bookkeeping must not reference work on the caller's stream), then capture there.
            if self._side is None:
                self._side = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            if self.calls < self.warmup:
                self.calls += 1
                with torch.cuda.stream(self._side):
                    loss = self._step_body().detach()
                cur.wait_stream(self._side)
                return loss
            from . import lib as _lib
            # Two graphs over one memory pool: forward + loss | backward + optimizer.  The loss is final when the first
            # one has run, so a loop can read it back (loss_value) while the backward is still executing and launch the
            # next step behind it — the GPU never waits for the host's per-step read-back.
            self.graph, self.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            step_before = self.opt._step
            n0 = _lib.launch_count()
            # thread_local: NCCL's watchdog thread may touch the CUDA runtime while this thread captures
            with torch.cuda.graph(self.graph, stream=self._side, capture_error_mode="thread_local"):
                loss, self._carry = self._fwd_body()
                self.loss = loss.detach()
            with torch.cuda.graph(self.graph_b, pool=self.graph.pool(), stream=self._side, capture_error_mode="thread_local"):
                self._bwd_body(self._carry)
            self.launches_per_step = _lib.launch_count() - n0   # kernels of ours inside one replay of both
            self.opt._step = step_before              # capture does not execute; the replay below is the real step
            self._loss_ready = torch.cuda.Event()
        self.graph.replay()
        self._loss_ready.record()
        self.graph_b.replay()
        self.opt._step += 1
        flat = self.model._flat
        flat._gather_pending = flat.world > 1    # the replayed step ended with an un-gathered optimizer update
        return self.loss
