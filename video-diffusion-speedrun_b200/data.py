"""Latent input pipeline: the step *before* the hot path.

The reference stacks per-sample latents in a DataLoader collate (``utils.py:18-35``) and moves the batch with a
synchronous ``images_vae.to(device).to(torch.bfloat16)`` at the top of every step (``train.py:73``); every rank builds
the same DataLoader with ``shuffle=True`` and no sampler, so ranks draw independently.  Here:

* ``shard_indices``  — a DistributedSampler-equivalent: one seeded permutation per epoch, strided by rank, padded or
  truncated so every rank gets the same number of samples (pure Python/torch-CPU, unit-tested without a GPU).
* ``DevicePrefetcher`` — pinned, double-buffered host->device copies on a dedicated copy stream: batch i+1 is in flight
  (H2D + the bf16 cast of ``train.py:73``) while step i computes; ``next()`` only makes the compute stream wait on the
  copy's event.  The dataset itself (HF ``fal/cosmos-openvid-1m`` rows, ``torch.load`` of ``serialized_latent``) is out
  of scope — any iterable of ``{"latent": [B,16,T,H,W] tensor, ...}`` host batches works.
"""
import torch


def shard_indices(n, rank, world, seed=0, epoch=0, shuffle=True, drop_last=False):
    """Indices of this rank's samples for one epoch.  Matches torch.utils.data.DistributedSampler's contract:
    same permutation on every rank (seed + epoch), rank r takes positions r, r+world, ...; without drop_last the
    tail is padded by wrapping around so all ranks see ceil(n / world) samples."""
    assert 0 <= rank < world and n >= 0
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    if drop_last:
        per = n // world
        idx = idx[:per * world]
    else:
        per = (n + world - 1) // world
        pad = per * world - n
        if pad > 0 and n > 0:
            idx = idx + (idx * ((pad + n - 1) // n))[:pad]
    return idx[rank:per * world:world]


def batches(indices, batch_size, drop_last=True):
    """Consecutive index lists of length batch_size (the DataLoader's batch sampler)."""
    out = [indices[i:i + batch_size] for i in range(0, len(indices), batch_size)]
    if drop_last and out and len(out[-1]) < batch_size:
        out.pop()
    return out


class DevicePrefetcher:
    """Iterate host batches with the next batch's H2D copy overlapped with the current step.

    ``it`` yields dicts (or tuples) whose tensor entries are CPU tensors; non-tensor entries (prompts) pass through.
    Floating-point tensors are delivered in ``dtype`` (bf16: the cast of train.py:73 happens on the device, after a
    copy of the source bytes — fp16/bf16 sources move half the bytes of an fp32-then-cast path).
    ``depth`` pinned staging slots are kept per tensor entry (2 = double buffering).
    """

    def __init__(self, it, device="cuda", dtype=torch.bfloat16, depth=2):
        self.it = iter(it)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("DevicePrefetcher needs a CUDA device (no CPU path)")
        self.dtype = dtype
        self.depth = max(1, int(depth))
        self.stream = torch.cuda.Stream(device=self.dev)
        self.queue = []          # [(batch_on_device, event, pinned_refs)]
        self.h2d_bytes = 0
        self._done = False
        for _ in range(self.depth):
            self._issue()

    def _pin(self, t):
        return t if t.is_pinned() else t.pin_memory()

    def _move(self, v, pinned_refs):
        if not torch.is_tensor(v):
            return v
        src = self._pin(v.contiguous())
        pinned_refs.append(src)
        self.h2d_bytes += src.numel() * src.element_size()
        d = src.to(self.dev, non_blocking=True)
        if d.is_floating_point() and d.dtype != self.dtype:
            d = d.to(self.dtype)
        return d

    def _issue(self):
        if self._done:
            return
        try:
            host = next(self.it)
        except StopIteration:
            self._done = True
            return
        refs = []
        with torch.cuda.stream(self.stream):
            if isinstance(host, dict):
                dev = {k: self._move(v, refs) for k, v in host.items()}
            else:
                dev = type(host)(self._move(v, refs) for v in host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.queue.append((dev, ev, refs))

    def __iter__(self):
        return self

    def __next__(self):
        if not self.queue:
            raise StopIteration
        dev, ev, _refs = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        for v in (dev.values() if isinstance(dev, dict) else dev):
            if torch.is_tensor(v):
                v.record_stream(cur)     # allocated on the copy stream, consumed on the compute stream
        self._issue()
        return dev
