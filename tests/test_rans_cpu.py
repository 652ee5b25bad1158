"""The native rANS coder (csrc/rans.cpp behind masic_b200/rans.py) writes the byte format of the reference's
`compressai.ans` extension: known-answer byte strings generated from the compiled reference extension
(tests/golden/make_golden_rans.py), a live cross-check against that extension where oracle/_ref holds it, the list
API and the buffer API agree, and stale-table protection of the entropy models."""
import hashlib
import json

import numpy as np
import pytest
import torch

from masic_b200 import rans


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.loads((golden_dir / "rans_kat.json").read_text())["cases"]


def test_known_answer_byte_strings(kat):
    for c in kat:
        sym, idx = np.asarray(c["symbols"], np.int32), np.asarray(c["indexes"], np.int32)
        enc = rans.RansEncoder().encode_with_indexes(c["symbols"], c["indexes"], c["cdfs"], c["sizes"], c["offsets"])
        assert len(enc) == c["n_bytes"] and hashlib.sha256(enc).hexdigest() == c["sha256"]
        if c["hex"]:
            assert enc.hex() == c["hex"]
        # buffer API: int32 arrays + a TableSet, no lists
        t = rans.TableSet(np.asarray(c["cdfs"], np.int32), c["sizes"], c["offsets"])
        assert rans.RansEncoder().encode_with_indexes(sym, idx, t) == enc
        # buffered pushes concatenate
        b = rans.BufferedRansEncoder()
        b.encode_with_indexes(sym[:7], idx[:7], t)
        b.encode_with_indexes(sym[7:], idx[7:], t)
        assert b.flush() == enc
        d = rans.RansDecoder()
        assert d.decode_with_indexes(enc, c["indexes"], c["cdfs"], c["sizes"], c["offsets"]) == c["symbols"]
        assert np.array_equal(d.decode_with_indexes_array(enc, idx, t), sym)
        d.set_stream(enc)                                    # streaming decode in two calls
        first = d.decode_stream(idx[:11], t)
        rest = d.decode_stream(idx[11:], t)
        assert first + rest == c["symbols"]


def test_live_cross_check_against_the_reference_extension():
    from oracle import refimport
    try:
        ref = refimport.load_ref_ext("ans")
    except ImportError:
        pytest.skip("oracle/_ref/ans not built (needs /root/reference)")
    rng = np.random.default_rng(7)
    for trial in range(20):
        n_tables = int(rng.integers(1, 40))
        ln = int(rng.integers(2, 60))
        f = rng.integers(1, 1000, (n_tables, ln))
        f = np.maximum(1, f * (65536 - ln) // f.sum(1, keepdims=True))
        f[:, 0] += 65536 - f.sum(1)
        cdf = np.concatenate([np.zeros((n_tables, 1), np.int64), np.cumsum(f, 1)], 1).astype(np.int32)
        sizes = np.full(n_tables, ln + 1, np.int32)
        offs = -rng.integers(0, ln, n_tables).astype(np.int32)
        n = int(rng.integers(1, 3000))
        idx = rng.integers(0, n_tables, n).astype(np.int32)
        sym = np.round(rng.normal(0, [2, 20, 3000][trial % 3], n)).astype(np.int32)
        want = ref.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), sizes.tolist(), offs.tolist())
        t = rans.TableSet(cdf, sizes, offs)
        got = rans.RansEncoder().encode_with_indexes(sym, idx, t)
        assert got == want
        assert ref.RansDecoder().decode_with_indexes(got, idx.tolist(), cdf.tolist(), sizes.tolist(), offs.tolist()) == sym.tolist()
        assert np.array_equal(rans.RansDecoder().decode_with_indexes_array(want, idx, t), sym)


def test_bad_arguments():
    from masic_b200._lib import MasicError
    t = rans.TableSet([[0, 30000, 65536]], [3], [0])
    with pytest.raises(MasicError):
        rans.RansEncoder().encode_with_indexes([0], [5], t)          # table index out of range
    with pytest.raises(ValueError):
        rans.RansEncoder().encode_with_indexes([0, 1], [0], t)
    with pytest.raises(MasicError):
        rans.RansDecoder().set_stream(b"\x00\x01")                   # shorter than the final state
    with pytest.raises(ValueError):
        rans.RansDecoder().decode_stream([0], t)                     # no stream


def test_entropy_model_table_cache_is_dropped_on_every_table_change():
    """compress -> update(force) -> update(force) -> compress must never code with stale tables."""
    from masic_b200.entropy_models import EntropyBottleneck, GaussianConditional
    torch.manual_seed(0)
    eb = EntropyBottleneck(8)
    eb.update()
    t0 = eb._coder_tables()
    assert eb._coder_tables() is t0                                   # cached between calls
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= 7.0
        eb.quantiles[:, 0, 2] += 5.0
    eb.update(force=True)
    t1 = eb._coder_tables()
    assert t1 is not t0 and t1[0].pitch != t0[0].pitch
    eb.update(force=True)
    assert eb._coder_tables() is not t1
    # load_state_dict path (buffers copied in place)
    eb2 = EntropyBottleneck(8)
    eb2.update()
    c0 = eb2._coder_tables()
    sd = eb.state_dict()
    for k in ("_offset", "_quantized_cdf", "_cdf_length"):
        eb2._buffers[k] = torch.zeros(sd[k].shape, dtype=torch.int32)      # bypasses __setattr__ on purpose
    eb2.load_state_dict(sd)
    c1 = eb2._coder_tables()
    assert c1 is not c0 and np.array_equal(c1[0].cdfs, eb._quantized_cdf.numpy())
    gc = GaussianConditional([0.11, 0.5, 1.0, 4.0])
    gc.update()
    g0 = gc._coder_tables()
    gc.update_scale_table([0.2, 0.7, 2.0], force=True)
    assert gc._coder_tables() is not g0 and gc._coder_tables()[0].n_tables == 3
