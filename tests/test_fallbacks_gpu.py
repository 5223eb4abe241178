"""The non-TMA fallbacks of the round-2 output paths stay correct.  Every path that hands a staged tile to the TMA unit
(dK / dV / O of the attention kernels, the wgrad accumulate, the fused QKV epilogue) keeps its per-thread fallback for
buffers a tensor map cannot describe; the tuning switches that force those fallbacks are exercised here in a child
process (the switches are read once per process) over the kernel and end-to-end parity tests that cover them."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_holds_with_every_tma_output_path_switched_off():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    env = dict(os.environ, VDS_BWD2_TMA_OUT="0", VDS_BWD_TMA_OUT="0", VDS_GEMM_TMA_RED="0", VDS_FUSE_QKV_ROPE="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu",
                        os.path.join(ROOT, "tests", "test_kernels_gpu.py"), os.path.join(ROOT, "tests", "test_gemm_gpu.py"),
                        os.path.join(ROOT, "tests", "test_model_gpu.py"),
                        "-k", "attn_fwd_bwd or wgrad or golden_case or full_sequence"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
    assert " passed" in r.stdout
