"""Generates the golden fixtures in this directory from the UNMODIFIED reference (oracle/_ref,
built from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Fixtures (committed, small):
  edlib_kat.json      edlibAlign(HW, PATH) known answers: SURVEY.md §3.5 cases + seeded random cases
  perread_<name>.npz  inputs (packed reads, names, parameters) and the outputs of the reference's
                      own TGSFilterTask::filter_sequence worker body: emitted records, DropInfo,
                      quality histograms and every QC table
  cli_<name>.json     tiny FASTQ run through the reference CLI: stdout records + stderr INFO lines
  prepass.npz         read ends + CheckBaseContent / adapterSearch outcome of the reference
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_lib  # noqa: E402
from tgsfilter_b200 import synth  # noqa: E402
from tgsfilter_b200.params import ADAPTER_LIB, FilterParams, rev_comp  # noqa: E402

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def rnd(rng, n, alpha=b"ACGT"):
    a = np.frombuffer(alpha, dtype=np.uint8)
    return a[rng.integers(0, len(a), n)].tobytes()


def edlib_cases():
    rng = np.random.default_rng(20261017)
    cases = [(b"ACGT", b"TTAGCTTT", 3), (b"ACGTACGT", b"ACGACGTTTTTACGTACG", 6),
             (b"ACGT", b"acgtACGN", 1), (b"AAAA", b"CCCCCCCC", 3), (b"AAAA", b"CCCCCCCC", 4),
             (ADAPTER_LIB[0], b"GG" + ADAPTER_LIB[0][:20] + b"T" + ADAPTER_LIB[0][20:] + b"ACGTAC", 11),
             (ADAPTER_LIB[8], ADAPTER_LIB[8][3:] * 2, 47)]
    for i in range(400):
        ql = int(rng.integers(1, 130)) if i % 5 else int(rng.integers(130, 257))
        alpha = [b"ACGT", b"AC", b"A", b"ACGTN", b"ACGTacgt"][int(rng.integers(0, 5))]
        q = rnd(rng, ql, alpha)
        mode = int(rng.integers(0, 4))
        if mode == 0:
            t = rnd(rng, int(rng.integers(1, 300)), alpha)
        elif mode == 1:
            t = rnd(rng, int(rng.integers(0, 80)), alpha) + synth.mutate(q, float(rng.random() * 0.3), rng) \
                + rnd(rng, int(rng.integers(0, 80)), alpha)
        elif mode == 2:
            t = rnd(rng, int(rng.integers(0, 40)), alpha) + synth.mutate(q, 0.1, rng) \
                + rnd(rng, int(rng.integers(0, 60)), alpha) + synth.mutate(q, 0.1, rng)
        else:
            t = (q[:max(1, ql // 3)] * 12)[:int(rng.integers(1, 300))]
        if not t:
            t = b"A"
        k = [-1, ql, ql + 5, max(0, ql - 3), int(ql * 0.1) + 1, int(rng.integers(0, ql + 1)),
             max(0, ql - 34)][int(rng.integers(0, 7))]
        cases.append((q, t, k))
    return cases


def small_reads(seed, n, lo, hi, adapters, fasta=False, weird=False):
    """Short reads with planted adapters at both ends / middle, some lower case and N."""
    rng = np.random.default_rng(seed)
    seqs, quals, names = [], [], []
    for i in range(n):
        L = int(rng.integers(lo, hi))
        s = bytearray(rnd(rng, L))
        meanq = rng.normal(22, 6)
        q = np.clip(np.rint(rng.normal(meanq, 5, L)), 1, 60).astype(np.uint8) + 33
        u = rng.random()
        ad = adapters[int(rng.integers(0, len(adapters)))]
        if u < 0.25 and L > 2 * len(ad):
            m = synth.mutate(ad, 0.06, rng)
            lead = int(rng.integers(0, 12))
            s[lead:lead + len(m)] = m
        elif u < 0.45 and L > 2 * len(ad):
            m = synth.mutate(ad, 0.06, rng)
            s[L - len(m) - int(rng.integers(0, 8)):][:len(m)] = m
            s = s[:L]
        elif u < 0.70 and L > 500:
            m = synth.mutate(ad, 0.04, rng)
            pos = int(rng.integers(200, L - 200))
            s[pos:pos + len(m)] = m
            if u < 0.50:
                m2 = synth.mutate(ad, 0.04, rng)
                pos2 = int(rng.integers(200, L - 200))
                s[pos2:pos2 + len(m2)] = m2
        s = bytes(s[:L]) if len(s) >= L else bytes(s) + rnd(rng, L - len(s))
        if weird and i % 3 == 0:
            arr = np.frombuffer(s, dtype=np.uint8).copy()
            idx = rng.integers(0, L, L // 20)
            arr[idx] = np.frombuffer(b"NacgtnRY", dtype=np.uint8)[rng.integers(0, 8, len(idx))]
            s = arr.tobytes()
        seqs.append(s)
        quals.append(q.tobytes())
        names.append((b"r%d extra=%d" % (i, i * 7)) if i % 4 == 1 else (b"r%d" % i))
    return synth.pack_reads(seqs, None if fasta else quals, names)


def perread_cases():
    lib0 = [ADAPTER_LIB[0], ADAPTER_LIB[1]]
    ont = [ADAPTER_LIB[8], rev_comp(ADAPTER_LIB[8])]
    long_ad = (ADAPTER_LIB[12] + ADAPTER_LIB[10])[:100]  # 100 bp: two Myers words
    cases = {}
    p = FilterParams(min_len=100, min_q=15.0, head_trim=0, tail_trim=0, adapters=lib0).apply_read_type("hifi")
    cases["hifi"] = (p, small_reads(1, 60, 120, 2500, lib0), 1)
    p = FilterParams(min_len=100, min_q=10.0, head_trim=7, tail_trim=4, adapters=ont).apply_read_type("ont")
    cases["ont_trim"] = (p, small_reads(2, 60, 60, 3000, ont), 1)
    p = FilterParams(min_len=200, max_len=1500, min_q=12.0, max_q=30.0, head_trim=0, tail_trim=0,
                     discard=True, adapters=ont).apply_read_type("ont")
    cases["ont_discard"] = (p, small_reads(3, 60, 100, 3000, ont), 1)
    p = FilterParams(min_len=100, min_q=10.0, head_trim=0, tail_trim=0, kmer=11, min_repeat=3,
                     adapters=lib0).apply_read_type("clr")
    cases["clr_repeat"] = (p, small_reads(4, 50, 300, 2500, lib0), 1)
    p = FilterParams(min_len=100, min_q=8.0, head_trim=3, tail_trim=0, mid_match_len=20, extra_len=10,
                     end_len=80, bc_len=60, adapters=[long_ad, rev_comp(long_ad), ADAPTER_LIB[6]]).apply_read_type("ont")
    cases["multiword"] = (p, small_reads(5, 50, 150, 2500, [long_ad, ADAPTER_LIB[6]], weird=True), 0)
    p = FilterParams(min_len=100, head_trim=5, tail_trim=5, qtype=0, adapters=lib0).apply_read_type("hifi")
    cases["fasta"] = (p, small_reads(6, 40, 120, 2000, lib0, fasta=True, weird=True), 0)
    p = FilterParams(min_len=100, filter=False, only_qc=True, adapters=[])
    cases["qc_only"] = (p, small_reads(7, 40, 120, 2000, lib0, weird=True), 1)
    return cases


def main():
    assert ref_lib.available(), "reference build missing: make -C oracle ref"
    # --- edlib
    cases = edlib_cases()
    ref = ref_lib.edlib_batch(cases)
    js = [{"q": q.decode(), "t": t.decode(), "k": k, "d": d, "aln_len": al, "locs": locs}
          for (q, t, k), (d, al, locs) in zip(cases, ref)]
    with open(os.path.join(HERE, "edlib_kat.json"), "w") as f:
        json.dump(js, f, separators=(",", ":"))
    # --- per-read worker body
    for name, (p, batch, outfq) in perread_cases().items():
        out = ref_lib.perread(p, batch, outfq)
        rec_text = b"".join(r[0] for r in out["records"])
        rec_names = b"\n".join(r[1] for r in out["records"])
        rec_lens = np.array([r[2] for r in out["records"]], dtype=np.int32)
        pd = {f.name: getattr(p, f.name) for f in __import__("dataclasses").fields(p) if f.name != "adapters"}
        np.savez_compressed(
            os.path.join(HERE, f"perread_{name}.npz"),
            bases=batch.bases, quals=batch.quals if batch.quals is not None else np.zeros(0, np.uint8),
            has_qual=batch.quals is not None, offsets=batch.offsets,
            names=np.frombuffer(b"\n".join(batch.names), dtype=np.uint8),
            params=json.dumps(pd), adapters=np.frombuffer(b"\n".join(p.adapters), dtype=np.uint8),
            outfq=outfq, drop_info=out["drop_info"], raw_hist=out["raw_hist"], clean_hist=out["clean_hist"],
            rec_text=np.frombuffer(rec_text, dtype=np.uint8), rec_names=np.frombuffer(rec_names, dtype=np.uint8),
            rec_lens=rec_lens, **{t: out[t] for t in ref_lib.TABLES})
        print(name, "reads", batch.n_reads, "records", len(out["records"]), "drop", out["drop_info"].tolist())
    # --- CLI
    batch = synth.make_config(1, 40, max_len=4000)
    rc, recs, err, _html = ref_lib.run_cli(["-x", "hifi", "-t", "1"], batch.to_fastq())
    info = [l for l in err.splitlines() if l.startswith("INFO") and "written to" not in l]
    with open(os.path.join(HERE, "cli_hifi.json"), "w") as f:
        json.dump({"config": 1, "n_reads": 40, "max_len": 4000, "args": ["-x", "hifi", "-t", "1"], "rc": rc,
                   "stdout": recs.decode(), "info": info}, f)
    print("cli rc", rc, "\n".join(info))
    # --- pre-pass
    rng = np.random.default_rng(99)
    n, row = 400, 150
    e5 = ACGT[rng.integers(0, 4, (n, row))]
    e3 = ACGT[rng.integers(0, 4, (n, row))]
    e5[:, :9] = ACGT[(rng.random((n, 9)) < 0.7) * 0 + (rng.random((n, 9)) < 0.2) * 3]  # biased head
    ad = ADAPTER_LIB[8]
    for r in range(0, n, 2):
        m = np.frombuffer(synth.mutate(ad, 0.08, rng), dtype=np.uint8)
        o = int(rng.integers(0, 25))
        e5[r, o:o + len(m)] = m[:row - o]
    out = ref_lib.prepass(150, 150, 1.0, 0.9, e5, e3, ADAPTER_LIB)
    np.savez_compressed(os.path.join(HERE, "prepass.npz"), ends5p=e5, ends3p=e3, trim5p=out["trim5p"],
                        trim3p=out["trim3p"], adapter5p=np.frombuffer(out["adapter5p"], dtype=np.uint8),
                        adapter3p=np.frombuffer(out["adapter3p"], dtype=np.uint8), dep5p=out["dep5p"],
                        dep3p=out["dep3p"])
    print("prepass", out)


if __name__ == "__main__":
    main()
