"""Host-side logic of the parameter sharding (apply_fsdp replacement): layout math and, with a real
world_size-2 gloo process group on CPU, the all-gather / reduce-scatter plumbing and state_dict assembly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

CFG = dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=64, depth=3, num_heads=2, mlp_ratio=4.0,
           cross_attn_input_size=32, residual_v=True, train_bias_and_rms=True, use_rope=True)


def _shapes():
    import vds_b200  # noqa: F401
    from vds_b200.model import DiT
    torch.manual_seed(0)
    m = DiT(**CFG)
    return m, [(n, tuple(p.shape)) for n, p in m.named_parameters()]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_layout_covers_every_parameter_exactly_once(world):
    from vds_b200.shard import ALIGN, Layout
    m, shapes = _shapes()
    lay = Layout(shapes, CFG["depth"], world)
    assert len(lay.groups) == CFG["depth"] + 1 and lay.full_total == sum(lay.group_numel)
    for g, n in enumerate(lay.group_numel):
        assert n % (ALIGN * world) == 0
    total = 0
    for name, shape in shapes:
        g, off, numel, shp = lay.param[name]
        assert off % ALIGN == 0 and shp == shape
        pieces = [lay.shard_range(name, r) for r in range(world)]
        assert sum(p[2] for p in pieces) == numel                      # ranks tile the parameter
        pos = 0
        for s, po, ln in pieces:
            if ln:
                assert po == pos
                pos += ln
        total += numel
    assert total == sum(p.numel() for p in m.parameters())
    # AdamW chunk tables: disjoint, inside the shard, cover exactly the owned elements, right group ids
    group_of = {n: (i % 5) for i, (n, _) in enumerate(shapes)}
    for r in range(world):
        st, ln, gid = lay.adam_chunks(r, group_of, chunk=1000)
        covered = torch.zeros(lay.shard_total, dtype=torch.int32)
        for s, l, g in zip(st, ln, gid):
            covered[s:s + l] += 1
        assert covered.max().item() <= 1
        assert covered.sum().item() == sum(lay.shard_range(n, r)[2] for n, _ in shapes)
        for n, _ in shapes:
            s, _, l = lay.shard_range(n, r)
            hits = [g for s2, l2, g in zip(st, ln, gid) if s <= s2 < s + l]
            assert all(h == group_of[n] for h in hits)


@pytest.mark.parametrize("world,G", [(1, 3), (2, 2), (8, 1), (3, 2)])
def test_layout_context_kv_groups_are_contiguous_stacks(world, G):
    """ckv_group = G: the context_kv Linear of every block moves into groups of G consecutive blocks whose weights
    (then biases) sit back to back, so G blocks' worth are ONE [G*2h, Dc] GEMM operand; everything else is unchanged."""
    from vds_b200.shard import Layout
    m, shapes = _shapes()
    depth, h, Dc = CFG["depth"], CFG["hidden_size"], CFG["cross_attn_input_size"]
    lay = Layout(shapes, depth, world, ckv_group=G)
    n_ckv = (depth + G - 1) // G
    assert lay.n_ckv == n_ckv and lay.n_groups == depth + 1 + n_ckv
    seen = set()
    for i in range(depth):
        gi, j, nb = lay.ckv_of_block(i)
        assert gi == i // G and j == i % G and nb == min(G, depth - gi * G)
        (ws, rows, cols), bias = lay.ckv_ranges(gi)
        assert rows == nb * 2 * h and cols == Dc
        s, numel = lay.full_range(f"blocks.{i}.context_kv.weight")
        assert s == ws + j * 2 * h * Dc and numel == 2 * h * Dc                   # block i = rows [j*2h, (j+1)*2h) of the stack
        bs, bn = lay.full_range(f"blocks.{i}.context_kv.bias")
        assert bias is not None and bs == bias[0] + j * 2 * h and bias[1] == nb * 2 * h
        assert lay.param[f"blocks.{i}.qkv.weight"][0] == i                        # the rest of the block stays in its group
        seen.add(gi)
    assert seen == set(range(n_ckv))
    covered = torch.zeros(lay.full_total, dtype=torch.int32)
    for name, shape in shapes:
        s, numel = lay.full_range(name)
        covered[s:s + numel] += 1
    assert covered.max().item() == 1 and covered.sum().item() == sum(p.numel() for p in m.parameters())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vds_b200  # noqa: F401
        from vds_b200.model import DiT, apply_fsdp
        torch.manual_seed(0)
        m = DiT(**CFG)
        ref = {n: p.detach().clone() for n, p in m.named_parameters()}
        names = list(ref)
        m = apply_fsdp(m, torch.bfloat16, torch.float32)
        flat = m._flat
        assert flat.world == world and flat.rank == rank
        assert [n for n, _ in m.named_parameters()] == names               # names preserved (get_mup_setup relies on it)
        # 1. gathered bf16 compute parameters == bf16 image of the full parameters
        P = flat.compute_params()
        for n in names:
            assert torch.equal(P[n], ref[n].to(torch.bfloat16)), n
        # 2. local parameters are this rank's slice of the flat fp32 master
        for n, p in m.named_parameters():
            s, po, ln = flat.layout.shard_range(n, rank)
            assert p.numel() == ln and torch.equal(p.detach().flatten(), ref[n].flatten()[po:po + ln])
        # 3. fp32 reduce-scatter averages rank-dependent gradients
        flat.begin_backward()
        for n in names:
            flat.grad_views[n].copy_(ref[n] * (rank + 1))
        for g in range(flat.depth):
            flat.block_backward_done(g)
        for gi in range(flat.layout.n_ckv):       # grouped context_kv parameters live in their own groups
            flat.ckv_backward_done(gi)
        flat.end_backward()
        mean = sum(range(1, world + 1)) / world
        for n, p in m.named_parameters():
            if n == "blocks.0.lambda_param":
                assert p.grad is None
                continue
            s, po, ln = flat.layout.shard_range(n, rank)
            assert torch.allclose(p.grad.flatten(), ref[n].flatten()[po:po + ln] * mean, rtol=1e-6, atol=1e-7), n
        # 3b. a second backward without zero_grad accumulates into the reduced shard (gradient accumulation at W > 1)
        flat.begin_backward()
        for n in names:
            flat.grad_views[n].copy_(ref[n] * (rank + 1) * 0.5)
        for g in range(flat.depth):
            flat.block_backward_done(g)
        for gi in range(flat.layout.n_ckv):
            flat.ckv_backward_done(gi)
        flat.end_backward()
        for n, p in m.named_parameters():
            if n == "blocks.0.lambda_param":
                continue
            s, po, ln = flat.layout.shard_range(n, rank)
            assert torch.allclose(p.grad.flatten(), ref[n].flatten()[po:po + ln] * mean * 1.5, rtol=1e-6, atol=1e-7), n
        for p in m.parameters():
            p.grad = None
        # 4. state_dict() returns full reference-shaped tensors
        sd = m.state_dict()
        for n in names:
            assert sd[n].shape == ref[n].shape and torch.equal(sd[n].cpu(), ref[n]), n
        # 5. get_mup_setup works on the sharded module (full shapes from paramstatus)
        groups, settings = m.get_mup_setup(1e-3, 1e-1, ["patch_proj", "context_kv", "positional_embedding"])
        assert sum(len(g["params"]) for g in groups) == len(names)
        # 6. sharded save in the reference layout (all ranks call, DCP de-duplicates) and reload on rank 0
        from vds_b200.checkpoint import read_checkpoint, save_checkpoint
        ck = os.path.join(os.environ["VDS_TEST_TMP"], "ck")
        save_checkpoint(m, ck, skip_rope=True)
        plain = read_checkpoint(ck)
        for n in names:
            assert torch.equal(plain[n], ref[n]), n
        assert "rope.freqs_hwt_cos" not in plain
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_allgather_reducescatter_statedict(tmp_path):
    world = 2
    os.environ["VDS_TEST_TMP"] = str(tmp_path)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
