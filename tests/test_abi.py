"""The C-ABI library loads and exports every symbol include/vds_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vds_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vds_[a-z0-9_]+)\s*\(", src)))


def test_header_is_plain_c():
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "vds_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_exports_every_declared_symbol():
    import vds_b200  # noqa: F401
    from vds_b200 import lib
    if not os.path.exists(lib.SO_PATH):
        lib.build()
    L = lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vds_b200.h but not exported"
    assert L.vds_abi_version() == 2
    # every bound signature refers to a declared symbol and vice versa
    bound = set(lib._SIGNATURES) | {"vds_last_error", "vds_abi_version", "vds_launch_count", "vds_attn_bwd_tail_ws_bytes",
                                     "vds_attn_bwd_tail_plan"}
    assert bound == set(names), (bound ^ set(names))


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not fall back."""
    import pytest
    import torch
    import vds_b200  # noqa: F401
    from vds_b200.model import DiT
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = DiT(in_channels=16, hidden_size=256, depth=1, num_heads=2, cross_attn_input_size=64)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 16, 2, 4, 4), torch.zeros(1, 8, 64), torch.tensor([0.5]))


def test_epilogue_enum_matches_header():
    """The Python constants of the GEMM epilogues are the header's enum values (a silent mismatch would select the
    wrong fused epilogue)."""
    import os
    import re

    import vds_b200  # noqa: F401
    from vds_b200 import lib
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "vds_b200.h")).read()
    enum = dict(re.findall(r"(VDS_EPI_[A-Z0-9_]+) = (\d+)", hdr))
    assert len(enum) == 8
    for name, val in enum.items():
        assert getattr(lib, name.replace("VDS_", "")) == int(val), name
    err = dict(re.findall(r"(VDS_ERR_[A-Z]+) = (-\d+)", hdr))
    assert lib.ERR_UNSUPPORTED == int(err["VDS_ERR_UNSUPPORTED"])
