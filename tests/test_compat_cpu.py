"""The reference's own coremasic/mywork/MASIC.py, UNMODIFIED, must import and build its model on top
of masic_b200/compat (the `compressai`, `kornia`, `range_coder` names it imports).  Needs the
reference tree, so it runs in the build container only (skipped on the GPU box)."""
import hashlib
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/coremasic/mywork/MASIC.py")

_SCRIPT = r"""
import sys, json, hashlib, torch
sys.path[:0] = [%(root)r, %(root)r + '/masic_b200/compat', '/root/reference/coremasic/mywork']
import MASIC, compressai, kornia
assert MASIC.__file__.startswith('/root/reference/'), MASIC.__file__
assert '/masic_b200/compat/' in compressai.__file__ and '/masic_b200/compat/' in kornia.__file__
torch.manual_seed(0)
net = MASIC.HSIC(128, 192, 5).eval()
sd = net.state_dict()
lay = json.load(open(%(root)r + '/tests/golden/hsic_layout.json'))
got = {k: (list(v.shape), str(v.dtype).replace('torch.', '')) for k, v in sd.items()}
assert got == {k: (s, d) for k, s, d in lay['entries']}
for k, h in lay['param_sha256'].items():
    assert hashlib.sha256(sd[k].numpy().tobytes()).hexdigest() == h, k
net.update()
assert tuple(net.entropy_bottleneck1._quantized_cdf.shape) == (128, 23)
assert sum(p.numel() for p in net.parameters()) == 35048349
assert float(net.aux_loss()) > 0
from masic_b200._lib import MasicError
try:
    net(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64), torch.eye(3)[None])
    raise SystemExit('CPU forward must fail loudly')
except MasicError:
    pass
print('OK')
"""


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present (GPU box)")
def test_unmodified_masic_py_builds_on_compat_packages():
    out = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": str(ROOT)}], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:] + out.stderr[-3000:]
