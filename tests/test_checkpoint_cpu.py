"""Checkpoint I/O in the reference layout: DCP directory -> temp.pt -> load_state_dict(assign=True)."""
import torch

import vds_b200  # noqa: F401
from vds_b200.checkpoint import ROPE_KEYS, load_checkpoint, read_checkpoint, save_checkpoint
from vds_b200.model import DiT

CFG = dict(in_channels=16, patch_size=2, time_patch_size=2, hidden_size=64, depth=2, num_heads=4, mlp_ratio=4.0,
           cross_attn_input_size=32, residual_v=True, train_bias_and_rms=True, use_rope=True)


def _model(seed):
    torch.manual_seed(seed)
    return DiT(**CFG)


def test_roundtrip_through_dcp_and_meta_model(tmp_path):
    src = _model(3)
    keys = save_checkpoint(src, tmp_path / "ck")
    ref_sd = src.state_dict()
    assert keys == sorted(ref_sd)
    plain = read_checkpoint(tmp_path / "ck")
    assert (tmp_path / "ck" / "temp.pt").exists()              # the reference's conversion artefact
    assert set(plain) == set(ref_sd)
    with torch.device("meta"):                                  # sample.py:41-53
        dst = DiT(**CFG)
    status = load_checkpoint(dst, tmp_path / "ck", device="cpu", dtype=torch.float32, strict=True)
    assert not status.missing_keys and not status.unexpected_keys
    for k, v in ref_sd.items():
        assert torch.equal(dst.state_dict()[k], v), k
    assert not any(p.is_meta for p in dst.parameters())


def test_skip_rope_keeps_strict_loading(tmp_path):
    src = _model(5)
    keys = save_checkpoint(src, tmp_path / "ck", skip_rope=True)
    assert not any(k in keys for k in ROPE_KEYS)
    size = sum(f.stat().st_size for f in (tmp_path / "ck").iterdir())
    assert size < 4 * sum(p.numel() for p in src.parameters()) + (1 << 20)     # parameters only, no rope tables
    with torch.device("meta"):
        dst = DiT(**CFG)
    status = load_checkpoint(dst, tmp_path / "ck", strict=True)
    assert not status.missing_keys
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k


def test_bf16_load_like_sampling(tmp_path):
    src = _model(7)
    save_checkpoint(src, tmp_path / "ck", skip_rope=True)
    with torch.device("meta"):
        dst = DiT(**CFG)
    load_checkpoint(dst, tmp_path / "ck", dtype=torch.bfloat16)                 # sample.py:55-63
    assert all(p.dtype == torch.bfloat16 for p in dst.parameters())
    assert dst.rope.freqs_hwt_cos.dtype == torch.bfloat16
    w = "blocks.1.mlp.0.weight"
    assert torch.equal(dst.state_dict()[w], src.state_dict()[w].bfloat16())


def test_interchange_with_live_reference_if_present(tmp_path):
    """Build container only: a checkpoint written from the unmodified reference DiT loads into this implementation
    (strict) and vice versa — same keys, same shapes, same values."""
    import importlib.util
    import os

    import pytest
    ref_path = "/root/reference/model.py"
    if not os.path.exists(ref_path):
        pytest.skip("reference not mounted (GPU box)")
    import torch.distributed.checkpoint as dcp
    spec = importlib.util.spec_from_file_location("ref_model_ckpt", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(11)
    rmodel = ref.DiT(**CFG)
    dcp.save(rmodel.state_dict(), checkpoint_id=tmp_path / "from_ref")          # train.py:553-584
    with torch.device("meta"):
        ours = DiT(**CFG)
    st = load_checkpoint(ours, tmp_path / "from_ref", strict=True)
    assert not st.missing_keys and not st.unexpected_keys
    for k, v in rmodel.state_dict().items():
        assert torch.equal(ours.state_dict()[k], v), k
    # and back: ours -> DCP -> temp.pt -> reference model, the reference's own loading code path (train.py:299-312)
    save_checkpoint(ours, tmp_path / "from_ours")
    sd = read_checkpoint(tmp_path / "from_ours")
    with torch.device("meta"):
        r2 = ref.DiT(**CFG)
    st = r2.load_state_dict({k: v.clone() for k, v in sd.items()}, assign=True, strict=True)
    assert not st.missing_keys and not st.unexpected_keys
    for k, v in rmodel.state_dict().items():
        assert torch.equal(r2.state_dict()[k], v), k
