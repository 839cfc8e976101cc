"""Loader of the committed golden fixtures (tests/golden/, made by make_golden.py)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from tgsfilter_b200.params import FilterParams
from tgsfilter_b200.synth import ReadBatch

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLES = ("raw5p_cnt", "raw5p_qual", "raw3p_cnt", "raw3p_qual", "clean5p_cnt", "clean5p_qual",
          "clean3p_cnt", "clean3p_qual", "raw_bin_cnt", "raw_bin_qual", "clean_bin_cnt",
          "clean_bin_qual")


def perread_names():
    return sorted(os.path.basename(p)[len("perread_"):-4]
                  for p in glob.glob(os.path.join(HERE, "perread_*.npz")))


def load_perread(name):
    z = np.load(os.path.join(HERE, f"perread_{name}.npz"))
    pd = json.loads(str(z["params"]))
    adapters = [a for a in z["adapters"].tobytes().split(b"\n") if a]
    params = FilterParams(**pd, adapters=adapters)
    quals = z["quals"] if bool(z["has_qual"]) else None
    batch = ReadBatch(z["bases"], quals, z["offsets"].astype(np.uint64),
                      z["names"].tobytes().split(b"\n"))
    exp = {k: z[k] for k in ("drop_info", "raw_hist", "clean_hist") + TABLES}
    exp["rec_text"] = z["rec_text"].tobytes()
    exp["rec_names"] = z["rec_names"].tobytes()
    exp["rec_lens"] = z["rec_lens"]
    exp["outfq"] = int(z["outfq"])
    return params, batch, exp


def load_edlib():
    with open(os.path.join(HERE, "edlib_kat.json")) as f:
        js = json.load(f)
    return [(c["q"].encode(), c["t"].encode(), c["k"], c["d"], c["aln_len"],
             [tuple(x) for x in c["locs"]]) for c in js]


def check_against_golden(exp, layout, counters_flat, records):
    """Asserts that counters (flat uint64 block) and formatted records equal the fixture."""
    L = layout
    np.testing.assert_array_equal(counters_flat[L.drop_info:L.drop_info + 17], exp["drop_info"])
    np.testing.assert_array_equal(counters_flat[L.raw_hist:L.raw_hist + 256], exp["raw_hist"])
    np.testing.assert_array_equal(counters_flat[L.clean_hist:L.clean_hist + 256], exp["clean_hist"])
    for t in TABLES:
        rows_total = L.bc_len if ("5p" in t or "3p" in t) else L.max_bins
        off = getattr(L, t)
        mine = counters_flat[off:off + rows_total * 5].reshape(rows_total, 5)
        ref = exp[t]
        np.testing.assert_array_equal(mine[:ref.shape[0]], ref, err_msg=t)
        assert not mine[ref.shape[0]:].any(), f"{t}: non-zero rows beyond the reference's table"
    assert b"".join(r[0] for r in records) == exp["rec_text"]
    assert b"\n".join(r[1] for r in records) == exp["rec_names"]
    np.testing.assert_array_equal(np.array([r[2] for r in records], dtype=np.int32), exp["rec_lens"])
