"""oracle/udh.py (CPU restatement of the udh homography front-end) against the fixture generated from the UNMODIFIED
reference (coremasic/mywork/model.py `Net`, the chain of test2_real.py:201-211): tests/golden/make_golden_udh.py."""
import hashlib

import numpy as np
import torch

from oracle import udh as OU


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_udh_oracle_matches_reference_fixture(golden_dir):
    fx = np.load(golden_dir / "udh.npz")
    torch.manual_seed(0)
    net = OU.OracleUDHNet(128).eval()
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in fx["keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in fx["shapes"]]
    assert hashlib.sha256(sd["fc.5.weight"].numpy().tobytes()).hexdigest() == str(fx["sha_fc5"])
    assert hashlib.sha256(sd["cnn.0.layers.0.weight"].numpy().tobytes()).hexdigest() == str(fx["sha_cnn0"])
    a, b, corners = OU.synthetic_patches(2, seed=3)
    assert torch.equal(a, _t(fx["a"])) and torch.equal(corners, _t(fx["corners"]))
    delta = net(a, b)
    assert torch.allclose(delta, _t(fx["delta"]), rtol=1e-5, atol=1e-7)
    h = OU.homography_from_delta(corners, _t(fx["delta"]), (1216, 2176))
    assert torch.allclose(h, _t(fx["h_1216x2176"]), rtol=1e-6, atol=1e-9)
    h_get = OU.homography_from_delta(corners, _t(fx["delta"]), (256, 256), shift_corners=False)
    assert torch.allclose(h_get, _t(fx["h_get_h"]), rtol=1e-6, atol=1e-9)


def test_masic_b200_udh_net_has_the_reference_state_dict(golden_dir):
    from masic_b200.udh import Net
    fx = np.load(golden_dir / "udh.npz")
    torch.manual_seed(0)
    sd = Net(patch_size=128).state_dict()
    assert list(sd.keys()) == [str(k) for k in fx["keys"]]
    assert hashlib.sha256(sd["fc.5.weight"].numpy().tobytes()).hexdigest() == str(fx["sha_fc5"])      # same seeded init
