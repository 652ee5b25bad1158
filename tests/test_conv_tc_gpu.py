"""GPU parity of the tcgen05 implicit-GEMM conv kernel: every layer shape class MASIC uses
(strided 5x5, transposed 5x5 in 4-phase and sub-pixel form, masked 5x5, 3x3, 1x1 incl. the
k=1 ConvTranspose2d, fused GDN / IGDN, per-pixel scaling, channel-slice concatenation, batch)
against torch fp32 convolutions on the same bf16-rounded operands and against the library's
CUDA-core direct conv.  Tolerance: bf16 output rounding (2^-8 relative to the tensor's range)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


@pytest.fixture(scope="module")
def diag():
    assert torch.cuda.is_available()
    import conv_diag
    return conv_diag


def _names():
    import importlib.util
    spec = importlib.util.spec_from_file_location("conv_diag_names", Path(__file__).resolve().parents[1] / "tools" / "conv_diag.py")
    # only the CASES dict keys are needed at collection time; avoid touching CUDA here
    src = Path(spec.origin).read_text()
    import re
    return re.findall(r'^    "([a-z0-9_]+)": dict\(', src, flags=re.M)


@pytest.mark.parametrize("name", _names())
def test_conv_case(diag, name):
    assert diag.run_case(name, **diag.CASES[name])


def test_plan_rejects_unsupported(diag):
    from masic_b200 import _lib
    from masic_b200.convplan import ConvPlan, GDN_FWD
    dev = torch.device("cuda:0")
    x = torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device=dev)
    out = torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(64, 64, 1, 1, device=dev)
    with pytest.raises(_lib.MasicError):        # fused GDN needs c_out == n_tile == 128
        ConvPlan(ksize=1, x=x, c_in=64, weight=w, c_out=64, n_tile=64, out=out, gdn=GDN_FWD,
                 gdn_beta=torch.ones(64, device=dev), gdn_gamma=torch.eye(64, device=dev))
    with pytest.raises(_lib.MasicError):        # n_tile must divide into 16s
        ConvPlan(ksize=1, x=x, c_in=64, weight=w, c_out=64, n_tile=24, out=out)
