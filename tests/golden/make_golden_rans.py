"""Known-answer byte strings of the reference's rANS extension (compressai.ans, compiled by `make -C oracle ref`
from compressai/cpp_exts/rans/rans_interface.cpp + third_party/ryg_rans/rans64.h) -> tests/golden/rans_kat.json.

    make -C oracle ref && python tests/golden/make_golden_rans.py

Cases cover: a single table, per-symbol table indexes, values below the offset / beyond the table (bypass-coded,
including raw values that need more than 15 nibbles' worth of counting... i.e. long escapes), ragged tables, and
a buffered encoder fed in several pushes."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle import refimport  # noqa: E402


def tables(rng, n_tables, max_len):
    cdfs, sizes, offsets = [], [], []
    for _ in range(n_tables):
        ln = int(rng.integers(2, max_len + 1))                 # number of intervals incl. the escape interval
        f = rng.integers(1, 4000, ln).astype(np.float64)
        f = np.maximum(1, np.floor(f / f.sum() * (65536 - ln))).astype(np.int64)
        f[int(rng.integers(0, ln))] += 65536 - f.sum()
        c = np.concatenate([[0], np.cumsum(f)]).astype(np.int64)
        assert c[-1] == 65536 and (np.diff(c) > 0).all()
        row = np.zeros(max_len + 1, dtype=np.int64)
        row[:ln + 1] = c
        cdfs.append(row.tolist())
        sizes.append(ln + 1)
        offsets.append(int(-rng.integers(0, ln)))
    return cdfs, sizes, offsets


def main():
    ans = refimport.load_ref_ext("ans")
    rng = np.random.default_rng(20261018)
    cases = []
    for ci, (n_tables, max_len, n, spread) in enumerate([(1, 8, 64, 3), (4, 23, 500, 12), (16, 40, 4000, 30),
                                                        (64, 120, 20000, 400), (3, 6, 300, 70000), (128, 23, 9000, 9)]):
        cdfs, sizes, offsets = tables(rng, n_tables, max_len)
        idx = rng.integers(0, n_tables, n).astype(np.int32)
        sym = np.round(rng.normal(0, spread, n)).astype(np.int32)
        sym[:4] = [0, -1, 1, spread * 5]
        enc = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), cdfs, sizes, offsets)
        dec = ans.RansDecoder().decode_with_indexes(enc, idx.tolist(), cdfs, sizes, offsets)
        assert dec == sym.tolist()
        b = ans.BufferedRansEncoder()
        cut = n // 3
        b.encode_with_indexes(sym[:cut].tolist(), idx[:cut].tolist(), cdfs, sizes, offsets)
        b.encode_with_indexes(sym[cut:].tolist(), idx[cut:].tolist(), cdfs, sizes, offsets)
        assert b.flush() == enc
        cases.append({"cdfs": cdfs, "sizes": sizes, "offsets": offsets, "indexes": idx.tolist(), "symbols": sym.tolist(),
                      "n_bytes": len(enc), "sha256": hashlib.sha256(enc).hexdigest(),
                      "hex": enc.hex() if len(enc) <= 4096 else None})
    out = {"source": "compressai.ans built from the unmodified reference (oracle/Makefile: ref)", "cases": cases}
    (HERE / "rans_kat.json").write_text(json.dumps(out))
    print("wrote", HERE / "rans_kat.json", [c["n_bytes"] for c in cases])


if __name__ == "__main__":
    main()
