"""CPU-side checks of the host logic: state_dict layout / seeded init of the product model,
CDF-table construction (update()), error behaviour of the entropy-model surface."""
import hashlib
import json

import numpy as np
import pytest
import torch

from masic_b200 import _lib
from masic_b200.entropy_models import EntropyBottleneck, GaussianConditional, GaussianMixtureConditional_gf
from masic_b200.hsic import HSIC


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_hsic_state_dict_layout_and_seeded_init(golden_dir):
    lay = json.loads((golden_dir / "hsic_layout.json").read_text())
    torch.manual_seed(0)
    net = HSIC()
    sd = net.state_dict()
    got = {k: (list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()}
    want = {k: (s, d) for k, s, d in lay["entries"]}
    assert got == want and len(got) == 248
    for k, h in lay["param_sha256"].items():
        assert hashlib.sha256(sd[k].numpy().tobytes()).hexdigest() == h, k
    assert sum(p.numel() for p in net.parameters()) == lay["main_params"] == 35048349
    assert sum(p.numel() for p in net.aux_parameters()) == lay["aux_params"] == 15616


def test_state_dict_round_trip_with_updated_tables():
    torch.manual_seed(0)
    a = HSIC()
    a.update()
    sd = a.state_dict()
    assert sd["entropy_bottleneck1._quantized_cdf"].shape == (128, 23)
    b = HSIC()
    b.load_state_dict(sd)                      # empty (0,) CDF buffers are resized to the checkpoint's
    assert torch.equal(b.entropy_bottleneck2._quantized_cdf, a.entropy_bottleneck2._quantized_cdf)


def test_eb_update_tables_match_reference(golden_dir):
    fx = np.load(golden_dir / "eb.npz")
    eb = EntropyBottleneck(16)
    sd = {k[3:]: _t(fx[k]) for k in fx.files if k.startswith("sd/")}
    for k in ("_offset", "_quantized_cdf", "_cdf_length"):
        sd[k] = torch.IntTensor()
    eb.load_state_dict(sd)
    eb.update()
    assert torch.equal(eb._offset, _t(fx["sd/_offset"]))
    assert torch.equal(eb._cdf_length, _t(fx["sd/_cdf_length"]))
    assert torch.equal(eb._quantized_cdf, _t(fx["sd/_quantized_cdf"]))
    assert float(eb.loss()) == pytest.approx(float(fx["aux_loss"]), rel=1e-5)
    fx0 = np.load(golden_dir / "eb_init_seed0.npz")
    torch.manual_seed(0)
    eb0 = EntropyBottleneck(128)
    eb0.update()
    assert torch.equal(eb0._quantized_cdf, _t(fx0["cdf"])) and torch.equal(eb0._offset, _t(fx0["offset"]))


def test_gc_update_tables_match_reference(golden_dir):
    fx = np.load(golden_dir / "gc.npz")
    gc = GaussianConditional([float(v) for v in fx["scale_table"]])
    gc.update()
    assert torch.equal(gc._offset, _t(fx["offset"])) and torch.equal(gc._cdf_length, _t(fx["length"]))
    assert list(gc._quantized_cdf.shape) == [64, 3133]
    assert torch.equal(gc._quantized_cdf[fx["cdf_rows"].tolist()], _t(fx["cdf_sel"]))
    assert hashlib.sha256(gc._quantized_cdf.numpy().tobytes()).digest() == fx["cdf_sha256"].tobytes()


def test_reference_error_behaviour():
    eb = EntropyBottleneck(4)
    with pytest.raises(ValueError, match="Invalid quantization mode"):
        eb._quantize(torch.zeros(1), "bogus")
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        eb._check_cdf_size()
    with pytest.raises(ValueError, match="Invalid scale_table"):
        GaussianConditional([3.0, 1.0])
    with pytest.raises(ValueError, match="Invalid type for scale_table"):
        GaussianConditional(3.0)
    g = GaussianMixtureConditional_gf(K=5)
    assert g.scale_table.numel() == 0 and float(g.scale_bound) == pytest.approx(0.11)
    with pytest.raises(_lib.MasicError):                       # product path never computes on the CPU
        HSIC().eval()(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64), torch.eye(3)[None])


def test_wave_schedule_respects_the_mask_a_context():
    """masic_b200/bitstream.py: every position of wave t = w + 3h only depends (5x5 mask 'A', layers.py:52-78) on
    positions of earlier waves, and the schedule is a permutation of the raster order."""
    from masic_b200.bitstream import wave_permutation, wave_schedule
    for h16, w16 in ((1, 1), (3, 2), (8, 12), (19, 34)):
        waves = wave_schedule(h16, w16)
        perm = wave_permutation(h16, w16)
        assert sorted(perm.tolist()) == list(range(h16 * w16))
        tmap = np.empty((h16, w16), dtype=np.int64)
        for t, (hs, ws) in enumerate(waves):
            tmap[hs, ws] = t
        for h in range(h16):
            for w in range(w16):
                for dh in (-2, -1, 0):
                    for dw in (-2, -1, 0, 1, 2):
                        if dh == 0 and dw >= 0:
                            continue
                        hh, ww = h + dh, w + dw
                        if 0 <= hh < h16 and 0 <= ww < w16:
                            assert tmap[hh, ww] < tmap[h, w]
    assert len(wave_schedule(76, 136)) == 361 and max(a.size for a, _ in wave_schedule(76, 136)) == 46
