"""Mirror of the denoising loop of the reference ``sampling/sample.py`` (generate_image, lines 77-159) on the
CUDA DiT: shifted-time Euler steps on the flow, classifier-free guidance with a zeroed negative embedding, fp32
latent accumulator, bf16 model input.  T5 encoding, the Streamlit UI and the Cosmos decoder / mp4 writer are out of
scope: the caller passes ``prompt_embeds`` and gets the final fp32 latents back.
"""
import torch


def shift_time(t, alpha=8.0):
    """sample.py:131-134."""
    return t * alpha / (1 + (alpha - 1) * t)


@torch.no_grad()
def denoise(model, prompt_embeds, latent_shape=(1, 16, 16, 64, 64), inference_steps=50, cfg_scale=6.0, seed=42,
            device="cuda", dtype=torch.bfloat16, latents=None, on_step=None, cache_context_kv=True):
    """sample.py:92-146.  Returns acc_latents (fp32).  Two separate model calls per step (cond / uncond), each drawing
    its own RoPE offsets from the global CPU generator, exactly like the reference."""
    prompt_embeds = prompt_embeds.to(device=device, dtype=dtype)
    negative_embeds = torch.zeros_like(prompt_embeds)                       # sample.py:104
    if latents is None:
        generator = torch.Generator(device=device).manual_seed(seed)         # sample.py:108
        latents = torch.randn(latent_shape, device=device, dtype=dtype, generator=generator)
    acc_latents = latents.to(dtype=torch.float32)
    if cache_context_kv and hasattr(model, "context_kv_cache"):
        model.context_kv_cache(True)     # prompt / negative embeddings are constant over the loop
    for i in range(inference_steps, 0, -1):
        t = shift_time(i / inference_steps)
        t_next = shift_time((i - 1) / inference_steps)
        dt = t - t_next
        tt = torch.tensor([t] * latents.shape[0]).to(device, dtype)
        model_output = model(latents, prompt_embeds, tt)
        if cfg_scale > 1:
            uncond_output = model(latents, negative_embeds, tt)
            model_output = uncond_output + cfg_scale * (model_output - uncond_output)
        acc_latents = acc_latents + dt * model_output.to(dtype=torch.float32)
        latents = acc_latents.to(dtype=dtype)
        if on_step is not None:
            on_step(i, acc_latents)
    if cache_context_kv and hasattr(model, "context_kv_cache"):
        model.context_kv_cache(False)
    return acc_latents


class GraphedDenoiser:
    """one whole denoising step — conditional forward, unconditional forward, classifier-free
    guidance, fp32 Euler update of the accumulator and the bf16 refresh of the model input (sample.py:122-146) —
    captured ONCE as a CUDA graph and replayed per step.  At B=1 the forward is ~600 short kernels, i.e. launch-bound
    from Python; one graph launch per step removes that.

    What changes per step lives in device buffers refreshed before each replay: the timestep, dt and the RoPE start
    offsets, which are still drawn from the global CPU generator in the reference's order (h, w, t; cond call first,
    then the uncond call), so a graphed run consumes the RNG exactly like ``denoise`` and is bit-identical to it.
    ``context_kv(prompt)`` of every block is computed once, before capture (§8f n2), and reused by every replay.

    ``batch_cfg=True`` runs cond and uncond as one 2B-sample forward (better tensor-core tiles at B=1).  The two halves
    then share one RoPE draw per step (3 draws instead of 6), which is a deliberate, documented deviation from the
    reference's RNG stream; everything else is unchanged.
    """

    def __init__(self, model, prompt_embeds, latent_shape, cfg_scale=6.0, device="cuda", dtype=torch.bfloat16,
                 batch_cfg=False):
        from .. import engine as _engine
        self.model, self._engine = model, _engine
        self.dev = torch.device(device)
        self.cfg_scale = float(cfg_scale)
        self.use_cfg = self.cfg_scale > 1
        self.batch_cfg = bool(batch_cfg) and self.use_cfg
        B, C, T, H, W = latent_shape
        self.B = B
        self.thw = (T // model.time_patch_size, H // model.patch_size, W // model.patch_size)
        cond = prompt_embeds.to(device=self.dev, dtype=dtype).contiguous()
        self.cond = cond
        self.uncond = torch.zeros_like(cond)                                    # sample.py:104
        self.both = torch.cat([cond, self.uncond], 0) if self.batch_cfg else None
        self.latents = torch.zeros(latent_shape, device=self.dev, dtype=dtype)
        self.lat2 = torch.zeros((2 * B, C, T, H, W), device=self.dev, dtype=dtype) if self.batch_cfg else None
        self.acc = torch.zeros(latent_shape, device=self.dev, dtype=torch.float32)
        nt = 2 * B if self.batch_cfg else B
        self.tt = torch.zeros((nt,), device=self.dev, dtype=dtype)
        self.dt = torch.zeros((), device=self.dev, dtype=torch.float32)
        self.n_fwd = 1 if (self.batch_cfg or not self.use_cfg) else 2
        self.starts_dev = [torch.zeros(3, device=self.dev, dtype=torch.int32) for _ in range(self.n_fwd)]
        self.graph = None
        self.launches_per_step = 0
        self._side = torch.cuda.Stream(device=self.dev)
        self._P = None

    def _body(self):
        model, eng = self.model, self._engine
        P = self._P
        if self.batch_cfg:
            self.lat2[:self.B].copy_(self.latents)
            self.lat2[self.B:].copy_(self.latents)
            out2, _ = eng.forward(model, P, self.lat2, self.both, self.tt, save=False, rope_starts_dev=self.starts_dev[0])
            out, un = out2[:self.B], out2[self.B:]
            out = un + self.cfg_scale * (out - un)
        else:
            out, _ = eng.forward(model, P, self.latents, self.cond, self.tt, save=False, rope_starts_dev=self.starts_dev[0])
            if self.use_cfg:
                un, _ = eng.forward(model, P, self.latents, self.uncond, self.tt, save=False,
                                    rope_starts_dev=self.starts_dev[1])
                out = un + self.cfg_scale * (out - un)
        self.acc.add_(out.to(torch.float32) * self.dt)                            # sample.py:143 (fp32 accumulator)
        self.latents.copy_(self.acc)                                             # sample.py:144 (bf16 model input)

    def _refresh(self, t, dt):
        # Fresh pinned staging tensors every step: torch's caching host allocator does not hand a block out again
        # before the async copy that reads it has run, so the host may run ahead of the GPU by many steps.
        for k in range(self.n_fwd):     # cond call first, then uncond: the reference's draw order
            st, sh, sw = self._engine.draw_rope_starts(self.model.rope, self.thw)
            self.starts_dev[k].copy_(torch.tensor([st, sh, sw], dtype=torch.int32).pin_memory(), non_blocking=True)
        # the reference builds the timestep tensor in fp32 and casts it to the model dtype (sample.py:137)
        tt = torch.full((self.tt.numel(),), t, dtype=torch.float32).to(self.tt.dtype).pin_memory()
        self.tt.copy_(tt, non_blocking=True)
        self.dt.copy_(torch.tensor(dt, dtype=torch.float32).pin_memory(), non_blocking=True)

    @torch.no_grad()
    def run(self, latents, inference_steps=50, on_step=None):
        """Denoise `latents` ([B,C,T,H,W]); returns the fp32 accumulator (a fresh tensor)."""
        model = self.model
        self.latents.copy_(latents)
        self.acc.copy_(latents.to(torch.float32))
        had_cache = getattr(model, "_ckv_cache", None) is not None
        if not had_cache:
            model.context_kv_cache(True)
        if self._P is None:
            self._P = model._param_view()
        try:
            for i in range(inference_steps, 0, -1):
                t = shift_time(i / inference_steps)
                dt = t - shift_time((i - 1) / inference_steps)
                self._refresh(t, dt)
                if self.graph is None:
                    from .. import lib as _lib
                    cur = torch.cuda.current_stream(self.dev)
                    # the first step runs eagerly on the capture stream: it also fills the context_kv cache, so the
                    # captured graph holds no K=4096 GEMM; the remaining steps replay the graph
                    self._side.wait_stream(cur)
                    with torch.cuda.stream(self._side):
                        self._body()
                    cur.wait_stream(self._side)
                    torch.cuda.synchronize(self.dev)
                    # the captured graph bakes in the addresses of the cached context_kv tensors: keep them alive
                    self._ckv_keep = dict(model._ckv_cache) if model._ckv_cache is not None else None
                    if i > 1:
                        keep_lat, keep_acc = self.latents.clone(), self.acc.clone()
                        self.graph = torch.cuda.CUDAGraph()
                        n0 = _lib.launch_count()
                        with torch.cuda.graph(self.graph, stream=self._side):
                            self._body()
                        self.launches_per_step = _lib.launch_count() - n0
                        self.latents.copy_(keep_lat)    # capture does not execute, but keep the state explicit
                        self.acc.copy_(keep_acc)
                else:
                    self.graph.replay()
                if on_step is not None:
                    on_step(i, self.acc)
        finally:
            if not had_cache:
                model.context_kv_cache(False)
        return self.acc.clone()
