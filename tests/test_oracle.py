"""CPU suite: the oracle restatement against the committed golden fixtures (always) and against
the reference build under oracle/_ref (when present)."""
import numpy as np
import pytest

import golden_lib
import oracle_lib
import ref_lib
from tgsfilter_b200 import records, synth
from tgsfilter_b200.params import ADAPTER_LIB

needs_ref = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")


def test_oracle_edlib_golden():
    for q, t, k, d, alen, locs in golden_lib.load_edlib():
        res, mylocs = oracle_lib.align_hw(q, t, k)
        assert (res["edit_distance"], res["align_len"], mylocs) == (d, alen, locs), (q, t, k)


@pytest.mark.parametrize("name", golden_lib.perread_names())
def test_oracle_perread_golden(name):
    params, batch, exp = golden_lib.load_perread(name)
    reads, pieces, cnt = oracle_lib.run(params, batch)
    recs = records.format_records(batch, pieces, fastq=exp["outfq"] == 1)
    golden_lib.check_against_golden(exp, oracle_lib.layout(params), cnt, recs)


def test_oracle_prepass_golden():
    z = np.load(golden_lib.HERE + "/prepass.npz")
    e5, e3 = z["ends5p"], z["ends3p"]
    n = e5.shape[0]
    t5 = oracle_lib.base_content_trim(oracle_lib.base_content_counts(e5), n, 1.0)
    t3 = oracle_lib.base_content_trim(oracle_lib.base_content_counts(e3), n, 1.0)
    assert (t5, t3) == (int(z["trim5p"]), int(z["trim3p"]))
    maps5 = oracle_lib.adapter_search(e5, ADAPTER_LIB, 0.9)
    best = int(np.argmax(maps5))
    assert ADAPTER_LIB[best] == z["adapter5p"].tobytes()
    assert np.float32(maps5[best]) / np.float32(len(ADAPTER_LIB[best])) == np.float32(z["dep5p"])


@needs_ref
@pytest.mark.ref
def test_oracle_edlib_vs_reference_random():
    rng = np.random.default_rng(5)
    a = np.frombuffer(b"ACGT", dtype=np.uint8)
    cases = []
    for _ in range(1500):
        ql = int(rng.integers(1, 200))
        q = a[rng.integers(0, 4, ql)].tobytes()
        t = (a[rng.integers(0, 4, int(rng.integers(0, 120)))].tobytes()
             + synth.mutate(q, float(rng.random() * 0.25), rng)
             + a[rng.integers(0, 4, int(rng.integers(0, 120)))].tobytes()) or b"A"
        k = [-1, ql, max(0, ql - 3), int(ql * 0.1) + 1, max(0, ql - 34)][int(rng.integers(0, 5))]
        cases.append((q, t, k))
    ref = ref_lib.edlib_batch(cases)
    for c, r in zip(cases, ref):
        res, locs = oracle_lib.align_hw(*c)
        assert (res["edit_distance"], res["align_len"], locs) == r, c


@needs_ref
@pytest.mark.ref
@pytest.mark.parametrize("cfg,n", [(1, 120), (2, 100), (3, 25), (4, 200), (5, 80)])
def test_oracle_perread_vs_reference(cfg, n):
    batch = synth.make_config(cfg, n, max_len=150000)
    params = synth.config_params(cfg)
    if cfg == 5:
        params.min_repeat = 40
    ref = ref_lib.perread(params, batch, 1)
    reads, pieces, cnt = oracle_lib.run(params, batch)
    L = oracle_lib.layout(params)
    np.testing.assert_array_equal(ref["drop_info"], cnt[L.drop_info:L.drop_info + 17])
    np.testing.assert_array_equal(ref["raw_hist"], cnt[L.raw_hist:L.raw_hist + 256])
    np.testing.assert_array_equal(ref["clean_hist"], cnt[L.clean_hist:L.clean_hist + 256])
    for t in ref_lib.TABLES:
        rows_total = L.bc_len if ("5p" in t or "3p" in t) else L.max_bins
        off = getattr(L, t)
        mine = cnt[off:off + rows_total * 5].reshape(rows_total, 5)
        np.testing.assert_array_equal(mine[:ref[t].shape[0]], ref[t], err_msg=t)
    assert records.format_records(batch, pieces) == ref["records"]
