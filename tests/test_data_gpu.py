"""DevicePrefetcher: overlapped H2D of the next latent batch; values / order / dtype must be exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_delivers_exact_batches_in_order(cuda_dev):
    from vds_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [{"latent": torch.randn((2, 16, 2, 8, 8), generator=g), "prompt": [f"p{i}", "q"],
             "ctx": torch.randn((2, 8, 32), generator=g).bfloat16()} for i in range(5)]
    pf = DevicePrefetcher(iter(host), device=cuda_dev, depth=2)
    seen = 0
    for i, b in enumerate(pf):
        _ = b["latent"].float().sum()      # some compute on the current stream between batches, as a step would do
        assert b["latent"].dtype == torch.bfloat16 and b["latent"].is_cuda
        assert torch.equal(b["latent"].cpu(), host[i]["latent"].bfloat16())     # train.py:73 cast
        assert torch.equal(b["ctx"].cpu(), host[i]["ctx"])
        assert b["prompt"] == host[i]["prompt"]
        seen += 1
    assert seen == 5
    assert pf.h2d_bytes == sum(h["latent"].numel() * 4 + h["ctx"].numel() * 2 for h in host)


def test_prefetcher_tuple_batches(cuda_dev):
    from vds_b200.data import DevicePrefetcher
    host = [(torch.full((4,), float(i)), torch.arange(3) + i) for i in range(3)]
    out = list(DevicePrefetcher(host, device=cuda_dev, depth=3))
    assert [int(o[0][0].item()) for o in out] == [0, 1, 2]
    assert out[2][1].dtype == torch.int64 and out[2][1].tolist() == [2, 3, 4]
