"""N > 1 on real GPUs (NCCL): run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("graph", [0, 1])
@pytest.mark.parametrize("cfg", ["tiny", "debug"])
def test_sharded_steps_match_single_gpu(world, graph, cfg):
    """W sharded ranks (own all-gather / reduce-scatter; Python-issued or captured in ONE CUDA graph per rank) after
    2-4 optimizer steps hold the parameters of a single-GPU run on the averaged gradients."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, VDS_MGPU_GRAPH=str(graph), VDS_MGPU_CFG=cfg)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29517 + world + 10 * graph + 20 * (cfg == "debug")),
                        os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=240,
                       env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "OK" in r.stdout
