"""The ctypes stub INTEGRATION.md section C shows a maintainer of the reference (the binding for
GaussianMixtureConditional_gf.forward and GaussianConditional.build_indexes) is executed VERBATIM from the document
against the built library and compared with the package's own bindings: the documentation cannot drift from the ABI."""
import re
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _stub_source():
    text = (ROOT / "INTEGRATION.md").read_text()
    sec = text[text.index("## C. The ctypes stub"):]
    m = re.search(r"```python\n(.*?)```", sec, re.S)
    assert m, "INTEGRATION.md section C lost its python block"
    return m.group(1)


def test_documented_ctypes_stub_runs_against_the_library():
    from masic_b200 import _lib, ops
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    src = _stub_source().replace('ctypes.CDLL("libmasic_b200.so")', f'ctypes.CDLL("{_lib.LIB_PATH}")')
    ns = {}
    exec(compile(src, "INTEGRATION.md#C", "exec"), ns)          # noqa: S102  (the document's own code)
    torch.manual_seed(0)
    n, m, k, h, w = 2, 192, 5, 19, 34
    y = (torch.randn(n, m, h, w, device=dev) * 3).contiguous()
    scales = (torch.rand(n, m * k, h, w, device=dev) * 2 + 0.05).contiguous()
    means = torch.randn(n, m * k, h, w, device=dev).contiguous()
    weights = torch.softmax(torch.randn(n, k, m, h, w, device=dev), 1).reshape(n, k * m, h, w).contiguous()
    y_hat, lik = ns["gmm_forward"](y, scales, means, weights, k)
    torch.cuda.synchronize()
    y_hat2, lik2 = ops.gmm_likelihood(y, scales, means, weights)[:2]
    assert torch.equal(y_hat, y_hat2) and torch.equal(lik, lik2)
    table = torch.exp(torch.linspace(-2.2, 5.5, 64)).to(dev)
    s2 = (torch.rand(3, 7, 11, device=dev) * 40).contiguous()
    idx = ns["build_indexes"](s2, table)
    torch.cuda.synchronize()
    assert torch.equal(idx, ops.gc_build_indexes(s2, table))
