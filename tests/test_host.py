"""CPU suite for the host side: C-ABI surface (load + exported symbols, loud failure without a
GPU), record assembly, parameter mirror, pre-pass decisions, sharding and the world_size-2 counter
all-reduce (gloo)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib
from tgsfilter_b200 import _capi, prepass, records, shard, synth
from tgsfilter_b200.params import ADAPTER_LIB, FilterParams, rev_comp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---- C-ABI surface ----------------------------------------------------------------------------
def test_library_loads_and_exports_every_declared_symbol():
    lib = _capi.load()
    header = open(os.path.join(ROOT, "include", "tgsf.h")).read()
    declared = set(re.findall(r"\b(tgsf_[a-z_0-9]+)\s*\(", header))
    declared -= {"tgsf_make_layout", "tgsf_bins_for_len"}
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.tgsf_version()


def test_struct_layouts_match_the_header():
    assert C.sizeof(_capi.ReadResult) == 32 and np.dtype(_capi.READ_RESULT_DTYPE).itemsize == 32
    assert C.sizeof(_capi.Piece) == 32 and np.dtype(_capi.PIECE_DTYPE).itemsize == 32
    assert C.sizeof(_capi.AlignResult) == 32 and np.dtype(_capi.ALIGN_RESULT_DTYPE).itemsize == 32
    assert C.sizeof(_capi.CounterLayout) == 18 * 4


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_engine_fails_loudly_without_gpu():
    from tgsfilter_b200.engine import FilterEngine
    with pytest.raises(_capi.TgsfError) as ei:
        FilterEngine(synth.config_params(1))
    assert ei.value.code == _capi.TGSF_ERR_CUDA


def test_create_rejects_bad_parameters_before_touching_the_gpu():
    lib = _capi.load()
    ctx = C.c_void_p()
    p = FilterParams(adapters=[b"A" * 2049]).apply_read_type("ont")  # TGSF_MAX_ADAPTER_LEN is 2048
    cp, keep = p.to_c()
    assert lib.tgsf_create(0, C.byref(cp), C.byref(ctx)) == _capi.TGSF_ERR_INVALID
    assert b"adapter" in lib.tgsf_last_error()
    p = FilterParams(adapters=[ADAPTER_LIB[0]], kmer=40, min_repeat=5).apply_read_type("hifi")
    cp, keep = p.to_c()
    assert lib.tgsf_create(0, C.byref(cp), C.byref(ctx)) == _capi.TGSF_ERR_INVALID
    p = FilterParams(adapters=[ADAPTER_LIB[0]])  # similarities never defaulted
    cp, keep = p.to_c()
    assert lib.tgsf_create(0, C.byref(cp), C.byref(ctx)) == _capi.TGSF_ERR_INVALID
    assert lib.tgsf_collect(None, None, 0, None, 0, None) == _capi.TGSF_ERR_INVALID


def test_counter_layout_is_shared_with_the_oracle():
    p = FilterParams(bc_len=150, max_read_len=250000)
    L = oracle_lib.layout(p)
    assert L.max_bins == 250000 // 100 + 1 and L.bc_len == 150
    assert L.drop_info == 0 and L.raw_hist == 17 and L.clean_hist == 17 + 256
    assert L.n_u64 == L.clean_bin_qual + L.max_bins * 5
    assert L.raw5p_cnt % 8 == 0 and L.raw_bin_cnt % 8 == 0


# ---- records ----------------------------------------------------------------------------------
def test_new_seq_name_inserts_before_first_whitespace():
    assert records.new_seq_name(b"read1", 2) == b"read1:2"
    assert records.new_seq_name(b"read1 runid=7 ch=2", 3) == b"read1:3 runid=7 ch=2"
    assert records.new_seq_name(b"r\tx y", 2) == b"r:2\tx y"
    assert records.new_seq_name(b"", 2) == b":2"


def test_format_records_numbers_only_emitted_pieces():
    batch = synth.pack_reads([b"ACGTACGTAC", b"GGGGGCCCCC"], [b"IIIIIIIIII", b"5555555555"],
                             [b"a desc", b"b"])
    pieces = np.zeros(4, dtype=_capi.PIECE_DTYPE)
    pieces["read"] = [0, 0, 0, 1]
    pieces["start"] = [0, 3, 6, 2]
    pieces["len"] = [2, 2, 4, 5]
    pieces["status"] = [_capi.PIECE_EMIT, _capi.PIECE_LOWQ, _capi.PIECE_EMIT, _capi.PIECE_EMIT]
    recs = records.format_records(batch, pieces)
    assert [r[1] for r in recs] == [b"a desc", b"a:2 desc", b"b"]
    assert recs[1][0] == b"@a:2 desc\nGTAC\n+\nIIII\n"
    assert records.format_records(batch, pieces, fastq=False)[2][0] == b">b\nGGGCC\n"


# ---- params -----------------------------------------------------------------------------------
def test_rev_comp_follows_the_reference_table():
    assert rev_comp(b"ACGTNacgtRYKMxz") == b"NNKMRYacgtNACGT"
    assert rev_comp(ADAPTER_LIB[0]) == ADAPTER_LIB[1]
    assert rev_comp(ADAPTER_LIB[8]) == ADAPTER_LIB[9]


def test_read_type_defaults():
    p = FilterParams().apply_read_type("hifi")
    assert (np.float32(p.mid_sim), np.float32(p.end_sim)) == (np.float32(0.95), np.float32(0.9))
    p = FilterParams(end_sim=0.7).apply_read_type("ont")
    assert (np.float32(p.mid_sim), np.float32(p.end_sim)) == (np.float32(0.9), np.float32(0.7))
    assert FilterParams().end_match_len == 4  # constructor default, not the usage text's 15


# ---- pre-pass host logic ----------------------------------------------------------------------
def test_base_content_trim_matches_oracle_on_random_counts():
    rng = np.random.default_rng(4)
    for trial in range(40):
        n = int(rng.integers(50, 5000))
        check = int(rng.integers(100, 200))
        base = rng.multinomial(n, [0.25] * 4, size=check).astype(np.int32)
        if trial % 2:
            base[: int(rng.integers(1, 30))] += rng.integers(0, n // 10 + 2, 4).astype(np.int32)
        bias = float([1.0, 0.5, 2.0, 5.0][trial % 4])
        assert prepass.base_content_trim(base, n, bias) == oracle_lib.base_content_trim(base, n, bias)


def test_qtype_and_default_min_q():
    assert prepass.get_qtype(34, 73) == 33
    assert prepass.get_qtype(66, 104) == 33 and prepass.get_qtype(80, 104) == 64
    assert prepass.get_qtype(50, 130) == 33 and prepass.get_qtype(60, 130) == 64
    assert prepass.default_min_q(-1, 40, "hifi") == 20 and prepass.default_min_q(-1, 15, "hifi") == 0
    assert prepass.default_min_q(-1, 40, "clr") == 10 and prepass.default_min_q(7.5, 40, "ont") == 7.5


def test_adapter_selection_rules():
    z = np.zeros(22, dtype=np.int64)
    c = np.zeros((150, 4), dtype=np.int32)
    # nothing found -> read-type fallback
    r = prepass.resolve(c, c, z, z, n=100, end_bias=1.0, mid_sim=0.9, bc_len=150, read_type="ont")
    assert r.adapters == [ADAPTER_LIB[8], ADAPTER_LIB[9]] and r.adapter5p == b""
    r = prepass.resolve(c, c, z, z, n=100, end_bias=1.0, mid_sim=0.95, bc_len=150, read_type="hifi")
    assert r.adapters == [ADAPTER_LIB[0], ADAPTER_LIB[1]]
    # a strong 5' adapter suppresses a > 5x weaker 3' one (T.cpp:3086-3092)
    m5, m3 = z.copy(), z.copy()
    m5[8] = 50 * 1000
    m3[4] = 28 * 150
    r = prepass.resolve(c, c, m5, m3, n=100, end_bias=1.0, mid_sim=0.9, bc_len=150, read_type="ont")
    assert r.adapter5p == ADAPTER_LIB[8] and r.adapter3p == b""
    assert r.adapters == [ADAPTER_LIB[8], ADAPTER_LIB[9]]
    # depth below 2*minSim is ignored (T.cpp:1193)
    m5 = z.copy()
    m5[8] = 50
    r = prepass.resolve(c, c, m5, z, n=100, end_bias=1.0, mid_sim=0.9, bc_len=150, read_type="clr")
    assert r.adapter5p == b"" and r.adapters == [ADAPTER_LIB[0], ADAPTER_LIB[1]]
    # explicit trims win over the base-content scan
    r = prepass.resolve(c, c, z, z, n=100, end_bias=1.0, mid_sim=0.9, bc_len=150, read_type="ont",
                        head_trim=0, tail_trim=12)
    assert (r.trim5p, r.trim3p) == (0, 12)


def test_sample_ends_follows_the_prepass_reader():
    batch = synth.make_config(1, 40, max_len=3000)
    p = FilterParams()
    e5, e3, mn, mx, check = prepass.sample_ends(batch, p)
    assert check == 150 and e5.shape == e3.shape and e5.shape[1] == 150
    lens = np.diff(batch.offsets.astype(np.int64))
    keep = np.nonzero(lens >= 1000)[0]
    assert e5.shape[0] == len(keep)
    b0, _ = batch.read(int(keep[0]))
    assert e5[0].tobytes() == b0[:150].tobytes()
    assert e3[0].tobytes() == rev_comp(b0[-150:].tobytes())
    assert 33 <= mn <= mx <= 33 + 60


# ---- sharding + counter all-reduce ------------------------------------------------------------
def test_split_batches_covers_every_read_once():
    batch = synth.make_config(2, 300, max_len=50000)
    bs = shard.split_batches(batch.offsets, 1_000_000)
    assert bs[0][0] == 0 and bs[-1][1] == batch.n_reads
    assert all(a[1] == b[0] for a, b in zip(bs[:-1], bs[1:]))
    assert all(hi > lo for lo, hi in bs)
    sizes = [int(batch.offsets[hi] - batch.offsets[lo]) for lo, hi in bs[:-1]]
    assert all(s <= 1_000_000 + 50000 for s in sizes)
    dealt = sorted(x for r in range(3) for x in shard.rank_batches(bs, r, 3))
    assert [d[1:] for d in dealt] == bs
    # degenerate: one read larger than the target
    bs1 = shard.split_batches(np.array([0, 10, 5000, 5010], dtype=np.uint64), 100)
    assert bs1 == [(0, 1), (1, 2), (2, 3)]


def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = synth.make_config(1, 60, max_len=4000)
    params = synth.config_params(1)
    params.head_trim = 5
    params.max_read_len = 8000
    cnt = np.zeros(oracle_lib.layout(params).n_u64, dtype=np.uint64)
    for _, lo, hi in shard.rank_batches(shard.split_batches(batch.offsets, 40_000), rank, world):
        _, _, cnt = oracle_lib.run(params, batch.slice(lo, hi), cnt)
    flat = torch.from_numpy(cnt.view(np.int64))
    shard.allreduce_counters(flat)
    if rank == 0:
        ret.put(flat.numpy().view(np.uint64).copy())
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_counters_allreduce_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    merged = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = synth.make_config(1, 60, max_len=4000)
    params = synth.config_params(1)
    params.head_trim = 5
    params.max_read_len = 8000
    _, _, single = oracle_lib.run(params, batch)
    np.testing.assert_array_equal(merged, single)


def test_pack_bases_roundtrip_on_cpu():
    lib = _capi.load()
    rng = np.random.default_rng(2)
    for n in (0, 1, 3, 4, 5, 31, 32, 33, 63, 64, 65, 1000, 4099, 100003):  # 32-base SIMD groups + scalar tails
        b = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        if n > 10:
            b[rng.integers(0, n, 7)] = np.frombuffer(b"NacgtRY", dtype=np.uint8)
        if n > 100:
            b[[31, 32, 64, n - 1]] = ord("N")  # group edges
        packed = np.zeros((n + 3) // 4 + 1, dtype=np.uint8)
        pos = np.zeros(16, dtype=np.uint64)
        val = np.zeros(16, dtype=np.uint8)
        ne = C.c_uint64(0)
        rc = lib.tgsf_pack_bases(b.ctypes.data if n else None, n, packed.ctypes.data, pos.ctypes.data,
                                 val.ctypes.data, 16, C.byref(ne))
        assert rc == 0
        codes = (packed[np.arange(n) // 4] >> (2 * (np.arange(n) % 4))) & 3
        rec = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].copy()
        rec[pos[:ne.value].astype(np.int64)] = val[:ne.value]
        assert np.array_equal(rec, b)
    # capacity error reports the required size
    b = np.frombuffer(b"NNNNNNNN", dtype=np.uint8).copy()
    packed = np.zeros(4, dtype=np.uint8)
    ne = C.c_uint64(0)
    assert lib.tgsf_pack_bases(b.ctypes.data, 8, packed.ctypes.data, None, None, 0, C.byref(ne)) == _capi.TGSF_ERR_CAPACITY
    assert ne.value == 8


def _write_fastx(path, n, fastq, crlf, rng):
    nl = b"\r\n" if crlf else b"\n"
    with open(path, "wb") as f:
        for i in range(n):
            ln = int(rng.integers(1, 2000))
            s = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), ln).tobytes()
            # qualities 33..74 include '@' and '+', so quality lines may start with either
            q = rng.integers(33, 75, ln).astype(np.uint8).tobytes()
            f.write((b"@r%d x" % i + nl + s + nl + b"+" + nl + q + nl) if fastq else (b">r%d" % i + nl + s + nl))


@pytest.mark.parametrize("fastq", [1, 0])
def test_parallel_chunk_parser_equals_serial_reader(tmp_path, fastq):
    """src/pipeline.hpp: record-boundary sync + ordered delivery of the parallel ingest equal the
    serial FastxReader-rule parser (T.cpp:685-760) on files whose quality lines start with '@'."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "ingest_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", os.path.join(root, "tests", "cpp", "ingest_check.cpp"),
                    "-lz", "-o", exe], check=True)
    rng = np.random.default_rng(17 + fastq)
    for crlf in (False, True):
        path = str(tmp_path / ("in_%d.fx" % crlf))
        _write_fastx(path, 1500, bool(fastq), crlf, rng)
        for chunk, thr in ((700, 3), (40000, 2), (1 << 22, 4)):
            r = subprocess.run([exe, path, str(fastq), str(chunk), str(thr)], capture_output=True, text=True)
            assert r.returncode == 0, (crlf, chunk, thr, r.stdout, r.stderr)
            assert r.stdout.split()[0] == "1500"
    # gzip input (two members, different levels) through the host's own inflate and through zlib
    if fastq:
        import gzip
        raw = open(path, "rb").read()
        gz = str(tmp_path / "in.fq.gz")
        open(gz, "wb").write(gzip.compress(raw[:len(raw) // 2], 6) + gzip.compress(raw[len(raw) // 2:], 1))
        env = dict(os.environ, INGEST_ONLY="serial", INGEST_HASH="1")
        want = subprocess.run([exe, path, "1", "50000", "1"], capture_output=True, text=True, env=env).stdout.split()[:3]
        got = subprocess.run([exe, gz, "1", "50000", "1"], capture_output=True, text=True, env=env).stdout.split()[:3]
        got_zlib = subprocess.run([exe, gz, "1", "50000", "1"], capture_output=True, text=True,
                                  env=dict(env, TGSF_ZLIB_INFLATE="1")).stdout.split()[:3]
        assert want == got == got_zlib and want[0] == "1500"
        # BGZF (block sizes in the headers: groups of blocks are decoded in parallel), BGZF followed by a
        # plain gzip member, and a corrupted block
        import struct
        import zlib

        def bgzf(data, blk=65280):
            out = bytearray()
            for i in list(range(0, len(data), blk)) + [None]:
                chunk = b"" if i is None else data[i:i + blk]
                c = zlib.compressobj(6, zlib.DEFLATED, -15)
                d = c.compress(chunk) + c.flush()
                out += (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(d) + 25) + d +
                        struct.pack("<II", zlib.crc32(chunk), len(chunk)))
            return bytes(out)

        env4 = dict(env, TGSF_INFLATE_THREADS="4")
        z = bgzf(raw)
        assert gzip.decompress(z) == raw
        for name, blob in (("bgzf", z), ("mixed", bgzf(raw[:len(raw) // 2]) + gzip.compress(raw[len(raw) // 2:]))):
            pth = str(tmp_path / (name + ".fq.gz"))
            open(pth, "wb").write(blob)
            assert subprocess.run([exe, pth, "1", "50000", "1"], capture_output=True, text=True, env=env4).stdout.split()[:3] == want
        bad = bytearray(z)
        bad[len(bad) // 2] ^= 0x10
        pth = str(tmp_path / "bad.fq.gz")
        open(pth, "wb").write(bytes(bad))
        r = subprocess.run([exe, pth, "1", "50000", "1"], capture_output=True, text=True, env=env4)
        assert "BGZF input" in r.stderr and int(r.stdout.split()[0]) < 1500
    # unterminated last line and an empty file
    tail = str(tmp_path / "tail.fq")
    open(tail, "wb").write(b"@a\nACGT\n+\nIIII\n@b\nAC\n+\nII")
    assert subprocess.run([exe, tail, "1", "7", "2"], capture_output=True).returncode == 0
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").close()
    assert subprocess.run([exe, empty, "1", "64", "2"], capture_output=True).returncode == 0


def test_own_inflate_matches_zlib(tmp_path):
    """src/inflate.hpp (gzip input of the C++ host) against zlib: stored, fixed and dynamic blocks, every
    strategy, multi-member files with header fields and zero padding, chunk-boundary sizes, corrupt and
    truncated streams (both decoders must reject them)."""
    import gzip
    import io
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "inflate_check")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "inflate_check.cpp"), "-lz", "-o", exe],
                   check=True)
    rng = np.random.default_rng(3)
    fq = synth.make_config(2, 120, with_names=False).to_fastq()
    datasets = {
        "fastq": fq,
        "random": rng.integers(0, 256, 1_500_000, dtype=np.uint8).tobytes(),
        "runs": (b"A" * 100000 + b"CG" * 50000 + b"ACGTACG" * 30000 + bytes(rng.integers(65, 70, 1000, dtype=np.uint8))) * 6,
        "text": b"the quick brown fox jumps over the lazy dog. " * 20000,
        "empty": b"",
        "one": b"x",
    }

    def comp(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
        c = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
        return c.compress(data) + c.flush()

    path = str(tmp_path / "t.gz")

    def check(blob, what, read_sizes=("1048576",)):
        with open(path, "wb") as f:
            f.write(blob)
        for rs in read_sizes:
            r = subprocess.run([exe, path, rs], capture_output=True, text=True)
            assert r.returncode == 0, (what, rs, r.stdout, r.stderr)

    for name, d in datasets.items():
        variants = [("l%d" % lv, comp(d, lv)) for lv in (0, 1, 6, 9)]
        variants += [("fixed", comp(d, 6, zlib.Z_FIXED)), ("huff", comp(d, 6, zlib.Z_HUFFMAN_ONLY)),
                     ("rle", comp(d, 6, zlib.Z_RLE))]
        bio = io.BytesIO()
        for part in (d[:len(d) // 3], d[len(d) // 3:2 * len(d) // 3], b"", d[2 * len(d) // 3:]):
            with gzip.GzipFile(filename="some_name.fq", mode="wb", fileobj=bio, compresslevel=5) as g:
                g.write(part)
        variants.append(("multi", bio.getvalue() + b"\0\0\0\0"))
        for vn, z in variants:
            check(z, (name, vn), ("1048576",) if len(d) > 100000 else ("7", "1048576"))
    z = bytearray(comp(fq[:200000], 6))
    for pos in (50, 1000, len(z) // 2, len(z) - 5, len(z) - 9):
        zz = bytearray(z)
        zz[pos] ^= 0x55
        check(bytes(zz), ("corrupt", pos))
    check(bytes(z[:len(z) // 2]), "truncated")


def test_speculative_multi_member_gzip_equals_sequential_decode(tmp_path):
    """fastgz::MultiMemberReader (cat-ed .fastq.gz files): spans between guessed member headers are decoded
    in parallel and accepted only as an unbroken chain from offset 0; a header-like byte string inside a
    stored block, a corrupt member and trailing padding must give exactly the sequential decoder's bytes
    (falling back to streaming where the chain breaks)."""
    import gzip
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "mm_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(root, "tests", "cpp", "multimember_check.cpp"),
                    "-lz", "-o", exe], check=True)
    rng = np.random.default_rng(2)
    recs = []
    for i in range(2000):
        ln = int(rng.integers(100, 30000))
        recs.append(b"@r%d\n" % i + rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln).tobytes() + b"\n+\n" +
                    rng.integers(33, 75, ln).astype(np.uint8).tobytes() + b"\n")
    raw = b"".join(recs)  # ~60 MB -> ~30 MB compressed: several 8 MB spans
    step = len(raw) // 7 + 1
    multi = b"".join(gzip.compress(raw[i:i + step], 1) for i in range(0, len(raw), step))
    path = str(tmp_path / "mm.gz")

    def run(blob, threads=4):
        with open(path, "wb") as f:
            f.write(blob)
        r = subprocess.run([exe, path, str(threads)], capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout, r.stderr)
        w = r.stdout.split()
        return dict(size=int(w[1]), failed=int(w[4]), fallback=int(w[6]), multi=int(w[8]))

    r = run(multi)
    assert r == dict(size=len(raw), failed=0, fallback=0, multi=1)
    assert run(gzip.compress(raw[:20_000_000], 1))["multi"] == 0           # single member: streaming path is chosen
    tiny = b"".join(gzip.compress(raw[i:i + 20000], 1) for i in range(0, 12_000_000, 20000))
    assert run(tiny, 3) == dict(size=12_000_000, failed=0, fallback=0, multi=1)
    fake = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + b"junk" * 8        # looks like a member header
    c = zlib.compressobj(0, zlib.DEFLATED, 31)                             # stored blocks keep it verbatim
    m1 = c.compress(raw[:9_000_000] + fake + raw[9_000_000:12_000_000]) + c.flush()
    r = run(m1 + gzip.compress(raw[12_000_000:30_000_000], 1) + gzip.compress(raw[30_000_000:40_000_000], 1))
    assert r["fallback"] == 1 and r["failed"] == 0 and r["size"] == 40_000_000 + len(fake)
    bad = bytearray(multi)
    bad[len(multi) // 2] ^= 0x20
    r = run(bytes(bad))
    assert r["failed"] == 1 and r["fallback"] == 1
    assert run(multi + b"\0" * 100) == dict(size=len(raw), failed=0, fallback=0, multi=1)


def test_parallel_single_stream_inflate_equals_sequential_decode(tmp_path):
    """fastgz::SingleStreamReader (src/pinflate.hpp; one-member .fastq.gz as written by gzip or pigz): spans
    start at block boundaries found by search and are decoded against an unknown 32 KB window, then accepted
    only as an unbroken chain from the member's first block.  Every zlib level and strategy, streams with
    sync/full flushes (pigz's chunking), stored-only and fixed-only streams, binary and mixed data (search
    finds nothing: spans are decoded again from the proven boundary), trailing members, corrupt bytes,
    truncation and a damaged trailer must give exactly the sequential decoder's bytes and verdict."""
    import gzip
    import io
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pinflate_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(root, "tests", "cpp", "pinflate_check.cpp"),
                    "-lz", "-o", exe], check=True)
    rng = np.random.default_rng(5)
    recs = []
    for i in range(220):
        ln = int(rng.integers(100, 30000))
        recs.append(b"@r%d\n" % i + rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln).tobytes() + b"\n+\n" +
                    rng.integers(33, 75, ln).astype(np.uint8).tobytes() + b"\n")
    fq = b"".join(recs)  # ~6.5 MB -> ~3.5 MB compressed: ~50 spans of 64 KB
    datasets = {
        "fastq": fq,
        "random": rng.integers(0, 256, 700_000, dtype=np.uint8).tobytes(),
        "runs": (b"A" * 100000 + b"CG" * 50000 + b"ACGTACG" * 30000 + bytes(rng.integers(65, 70, 1000, dtype=np.uint8))) * 3,
        "mixed": fq[:1_500_000] + rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes() + fq[1_500_000:3_000_000],
        "empty": b"",
        "one": b"x",
    }

    def comp(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
        c = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
        return c.compress(data) + c.flush()

    def flushed(data, kind):
        c = zlib.compressobj(6, zlib.DEFLATED, 31)
        out = b"".join(c.compress(data[i:i + 131072]) + c.flush(kind) for i in range(0, len(data), 131072))
        return out + c.flush()

    path = str(tmp_path / "t.gz")

    def run(blob, what, threads=("1", "5"), span="65536", rs="1048576", failed=0):
        with open(path, "wb") as f:
            f.write(blob)
        res = None
        for t in threads:
            r = subprocess.run([exe, path, t, span, rs], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, (what, t, r.stdout, r.stderr)
            w = r.stdout.split()
            assert (int(w[3]), int(w[4])) == (failed, failed), (what, t, r.stdout)
            res = dict(size=int(w[1]), speculated=int(w[-5]), spans=int(w[-3]), repairs=int(w[-1]))
        return res

    for name, d in datasets.items():
        variants = [("l%d" % lv, comp(d, lv)) for lv in (0, 1, 6, 9)]
        variants += [("fixed", comp(d, 6, zlib.Z_FIXED)), ("huff", comp(d, 6, zlib.Z_HUFFMAN_ONLY)),
                     ("sync", flushed(d, zlib.Z_SYNC_FLUSH)), ("full", flushed(d, zlib.Z_FULL_FLUSH))]
        bio = io.BytesIO()
        for part in (d[:len(d) // 2], b"", d[len(d) // 2:]):
            with gzip.GzipFile(filename="some_name.fq", mode="wb", fileobj=bio, compresslevel=5) as g:
                g.write(part)
        variants.append(("multi", bio.getvalue() + b"\0\0\0\0"))
        for vn, z in variants:
            r = run(z, (name, vn))
            assert r["size"] == len(d)
            if name == "fastq" and vn in ("l1", "l6", "l9", "huff", "sync", "full"):
                # text in dynamic blocks: the search lines the spans up (a flush's empty stored block is decoded through)
                assert r["spans"] > 20 and r["repairs"] <= r["spans"] // 10, (vn, r)
                assert r["speculated"] >= r["spans"] // 2, (vn, r)  # decoded against the unknown window, not one by one
    # memory bound: a span that expands beyond the soft cap is closed at a block boundary and the spans behind it
    # are decoded from there in turn, the last one as often as it takes (TGSF_PINFLATE_SOFT_CAP is a test knob)
    os.environ["TGSF_PINFLATE_SOFT_CAP"] = "50000"
    try:
        for name in ("runs", "fastq", "mixed"):
            r = run(comp(datasets[name], 6), (name, "soft cap"))
            assert r["size"] == len(datasets[name])
            assert name == "runs" or r["repairs"] > 0  # ("runs" is a single block: nothing to close early)
    finally:
        del os.environ["TGSF_PINFLATE_SOFT_CAP"]
    z = bytearray(comp(fq[:3_000_000], 6))
    for pos in (50, 1000, len(z) // 2, len(z) - 200, len(z) - 9, len(z) - 5):  # the last two: CRC-32, ISIZE
        zz = bytearray(z)
        zz[pos] ^= 0x55
        run(bytes(zz), ("corrupt", pos), failed=1)
    run(bytes(z[:len(z) // 2]), "truncated", failed=1)
    run(bytes(z[:len(z) - 4]), "truncated trailer", failed=1)
    run(bytes(z), "odd read size", threads=("4",), rs="4099")
    run(bytes(z), "default and oversized span", threads=("4",), span="0")
    run(bytes(z), "default and oversized span", threads=("4",), span="16777216")


def test_bam_and_sam_ingest_equals_fastq(tmp_path, cpp_tool):
    """BAM / SAM input (read_bam, T.cpp:1872-1916) is parsed by src/pipeline.hpp without htslib: the records must
    equal those of the equivalent FASTQ — every record whatever its flags, bases through the reference's nibble
    table (lower case folded, ambiguity codes and '=' -> NUL, unknown -> N), quality + 33 (absent: 0xFF + 33 = ' '),
    zero-length records kept.  Where the compiled reference is available its htslib must accept the test files:
    the reference CLI gives the same records and INFO lines for the BAM, the SAM and the FASTQ."""
    import subprocess
    import bam_lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = cpp_tool("ingest_check")
    rng = np.random.default_rng(3)

    def digest(path, sambam, threads="4", chunk="3000000"):
        env = dict(os.environ, INGEST_ONLY="serial", INGEST_HASH="1", TGSF_INFLATE_THREADS=threads)
        if sambam:
            env["INGEST_SAMBAM"] = "1"
        r = subprocess.run([exe, path, "1", chunk, "1"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        return r.stdout.split()[:3]

    def write(name, blob):
        path = str(tmp_path / name)
        with open(path, "wb") as f:
            f.write(blob)
        return path

    recs, sam, fq = [], [b"@HD\tVN:1.6\n", b"@PG\tID:x\n"], []
    for i in range(1200):
        ln = int(rng.integers(1, 4)) if i % 300 == 7 else int(rng.integers(50, 30000))
        seq = "".join(rng.choice(list("ACGT"), ln))
        if i % 10 == 0 and ln > 20:
            s = list(seq)
            for p in rng.integers(0, ln, 5):
                s[int(p)] = "NacgtnRY=xM"[int(rng.integers(0, 11))]
            seq = "".join(s)
        qual = None if i % 400 == 3 else rng.integers(0, 60, ln).astype(np.uint8)
        name = "m64011_%d/ccs" % i
        recs.append(bam_lib.bam_record(name, seq, qual, flag=(4, 0, 16, 256, 2048)[i % 5], tags=b"RGZgrp\0" if i % 3 else b""))
        sam.append(bam_lib.sam_line(name, seq, qual, tags="RG:Z:grp" if i % 3 else ""))
        q33 = b" " * ln if qual is None else bytes(qual + 33)
        fq.append(b"@" + name.encode() + b"\n" + bam_lib.through_reference_table(seq) + b"\n+\n" + q33 + b"\n")
    bam = write("x.bam", bam_lib.bam_file(recs, refs=(("chr1", 1000), ("chrUn_with_a_long_name", 5))))
    want = digest(write("x.fq", b"".join(fq)), False)
    assert want[0] == "1200"
    assert digest(bam, True) == want and digest(bam, True, threads="1") == want and digest(bam, True, chunk="20000") == want
    assert digest(write("x.sam", b"".join(sam)), True) == want
    # zero-length records (SEQ '*') stay in the stream; a truncated BAM ends the input at the last whole record
    recs0 = [bam_lib.bam_record("a", "ACGT", [1, 2, 3, 4]), bam_lib.bam_record("empty", "", []), bam_lib.bam_record("b", "GG", None)]
    sam0 = bam_lib.sam_line("a", "ACGT", [1, 2, 3, 4]) + bam_lib.sam_line("empty", "", None) + bam_lib.sam_line("b", "GG", None)
    d0 = digest(write("z.bam", bam_lib.bam_file(recs0)), True)
    assert d0[:2] == ["3", "6"] and digest(write("z.sam", sam0), True) == d0
    import zlib
    raw = b"BAM\1" + (0).to_bytes(4, "little") * 2 + b"".join(recs0)
    assert digest(write("t.bam", bam_lib.bgzf(raw[:-3])), True)[:2] == ["2", "4"]
    assert digest(write("e.bam", bam_lib.bam_file([])), True)[:2] == ["0", "0"]
    import ref_lib
    if ref_lib.available():
        clean = synth.make_config(2, 60, max_len=20000, with_names=False).to_fastq()
        b2, s2 = bam_lib.from_fastq(clean)
        outs = [ref_lib.run_cli(["-x", "ont", "-t", "1"], blob, in_name=nm) for nm, blob in (("in.fq", clean), ("in.bam", b2), ("in.sam", s2))]
        info = [[l for l in o[2].splitlines() if l.startswith("INFO") and "written to" not in l] for o in outs]
        assert outs[0][0] == 0 and len(outs[0][1]) > 10000
        assert outs[1][1] == outs[0][1] and outs[2][1] == outs[0][1] and info[1] == info[0] and info[2] == info[0]
        assert digest(write("c.bam", b2), True) == digest(write("c.fq", clean), False) == digest(write("c.sam", s2), True)


def test_stream_parallel_parse_equals_single_parser(tmp_path, cpp_tool):
    """TGSF_STREAM_PARSE_THREADS (opt-in, src/pipeline.hpp stream_reader_main): the decoded bytes of gzip / BGZF input
    are cut into chunks at verified record starts and parsed by several threads; records, order and the behaviour
    at a malformed record must be those of the single parser thread (FASTQ incl. quality lines starting with '@',
    2-line FASTA, records longer than a chunk)."""
    import gzip
    import subprocess
    import bam_lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = cpp_tool("ingest_check")
    rng = np.random.default_rng(8)
    recs, fa = [], []
    for i in range(900):
        ln = int(rng.integers(1, 5)) if i % 100 == 3 else int(rng.integers(50, 20000))
        if i == 450:
            ln = 400000  # longer than a chunk
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), ln).tobytes()
        q = bytearray(rng.integers(33, 75, ln).astype(np.uint8).tobytes())
        if i % 7 == 0:
            q[0] = ord("@")
        recs.append(b"@r%d some comment\n" % i + seq + b"\n+\n" + bytes(q) + b"\n")
        fa.append(b">r%d\n" % i + seq + b"\n")
    fq = b"".join(recs)
    bad = b"".join(recs[:600]) + b"@broken\nACGT\n+\nII\n" + b"".join(recs[600:])

    def digest(path, fastq, stream, chunk):
        env = dict(os.environ, INGEST_ONLY="serial", INGEST_HASH="1", TGSF_INFLATE_THREADS="4", TGSF_PINFLATE_MIN_BYTES="100000")
        if stream:
            env["TGSF_STREAM_PARSE_THREADS"] = "3"
        r = subprocess.run([exe, path, "1" if fastq else "0", chunk, "1"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        return r.stdout.split()[:3]

    def write(name, blob):
        path = str(tmp_path / name)
        with open(path, "wb") as f:
            f.write(blob)
        return path

    for name, blob, fastq in (("a.fq.gz", gzip.compress(fq, 4), True), ("b.fq.gz", bam_lib.bgzf(fq), True),
                              ("c.fa.gz", gzip.compress(b"".join(fa), 4), False), ("d.fq.gz", gzip.compress(bad, 4), True)):
        path = write(name, blob)
        for chunk in ("100000", "3000000"):
            want = digest(path, fastq, False, chunk)
            assert digest(path, fastq, True, chunk) == want, (name, chunk)
        assert int(want[0]) == (600 if name == "d.fq.gz" else 900)


def test_reader_stops_mid_stream_and_restarts(tmp_path):
    """Two-pass mode of the CLI (pre-pass sample over the buffer budget): the reader thread is stopped after a few
    batches, drained, joined and started again from the beginning of the file.  With the parallel decoders behind it
    (single stream, BGZF) the first round must end promptly and the second must deliver the whole file."""
    import gzip
    import subprocess
    import bam_lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "restart_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(root, "tests", "cpp", "restart_check.cpp"),
                    "-lz", "-o", exe], check=True)
    fq = synth.make_config(2, 400, max_len=30000, with_names=False).to_fastq()
    n_reads, n_bases = fq.count(b"\n") // 4, sum(len(l) for l in fq.split(b"\n")[1::4])
    for name, blob in (("a.fq.gz", gzip.compress(fq, 1)), ("b.fq.gz", bam_lib.bgzf(fq))):
        path = str(tmp_path / name)
        with open(path, "wb") as f:
            f.write(blob)
        for extra in ({}, {"TGSF_STREAM_PARSE_THREADS": "2"}):
            env = dict(os.environ, TGSF_INFLATE_THREADS="5", TGSF_PINFLATE_MIN_BYTES="100000", **extra)
            r = subprocess.run([exe, path, "500000", "2"], capture_output=True, text=True, env=env, timeout=120)
            assert r.returncode == 0, r.stderr
            rounds = [l.split() for l in r.stdout.splitlines()]
            assert int(rounds[0][4]) < n_bases and int(rounds[0][-2]) >= 2
            assert (int(rounds[1][2]), int(rounds[1][4])) == (n_reads, n_bases)


def test_multi_member_fallback_decodes_the_rest_in_parallel(tmp_path, cpp_tool):
    """When MultiMemberReader's chain of guessed member starts breaks (here: a gzip-header-like byte string inside a
    stored block) the remainder goes to the parallel single-stream reader (src/pinflate.hpp registers itself as the
    fallback) instead of the streaming decoder: the records must equal those of the sequential decoder and of zlib."""
    import gzip
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = cpp_tool("ingest_check")
    rng = np.random.default_rng(12)
    recs = [b"@r%d\n" % i + rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 20000).tobytes() + b"\n+\n" +
            rng.integers(33, 75, 20000).astype(np.uint8).tobytes() + b"\n" for i in range(900)]
    fake = b"@x \x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03 y\nACGT\n+\nIIII\n"
    c = zlib.compressobj(0, zlib.DEFLATED, 31)  # stored blocks keep the fake header verbatim
    m1 = c.compress(b"".join(recs[:250]) + fake + b"".join(recs[250:300])) + c.flush()
    blob = m1 + gzip.compress(b"".join(recs[300:600]), 1) + gzip.compress(b"".join(recs[600:]), 1)
    path = str(tmp_path / "mm.fq.gz")
    with open(path, "wb") as f:
        f.write(blob)

    def digest(**extra):
        env = dict(os.environ, INGEST_ONLY="serial", INGEST_HASH="1", TGSF_INFLATE_THREADS="5", **extra)
        r = subprocess.run([exe, path, "1", "3000000", "1"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        return r.stdout.split()[:3]

    want = digest(TGSF_ZLIB_INFLATE="1")
    assert want[0] == "901"
    assert digest() == want and digest(TGSF_SERIAL_INFLATE="1") == want and digest(TGSF_STREAM_PARSE_THREADS="2") == want


def test_gzip_member_format_of_the_gpu_encoder_on_cpu(tmp_path):
    """tgsfilter_b200/csrc/gzenc_core.h (code lengths, canonical codes, dynamic block header — the serial half
    of the GPU deflate encoder) built for the host: members assembled from it must inflate with zlib to the
    records; length limiting must keep the codes complete."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "gzenc_check")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "gzenc_check.cpp"), "-lz", "-o", exe],
                   check=True)
    assert subprocess.run([exe, "--selftest"], capture_output=True).returncode == 0
    rng = np.random.default_rng(5)
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    q = rng.permutation(np.frombuffer(b"".join(bytes([40 + i]) * c for i, c in enumerate(fib)), dtype=np.uint8)).tobytes()
    recs = [b"@fib skewed qualities\n" + rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), len(q)).tobytes() + b"\n+\n" + q + b"\n",
            b"@a\nA\n+\nI\n", b"@homo\n" + b"T" * 5000 + b"\n+\n" + b"#" * 5000 + b"\n",
            b"@" + b"n" * 300 + b"\nACGTNacgtn\n+\n!~!~!~!~!~\n",
            b"@wide\n" + rng.choice(np.frombuffer(b"ACGTNRYKM", dtype=np.uint8), 20000).tobytes() + b"\n+\n" +
            rng.integers(33, 127, 20000).astype(np.uint8).tobytes() + b"\n"]
    odd = str(tmp_path / "odd.fq")
    open(odd, "wb").write(b"".join(recs))
    ont = str(tmp_path / "ont.fq")
    open(ont, "wb").write(synth.make_config(2, 60, with_names=False).to_fastq())
    for path in (odd, ont):
        for fasta in ("0", "1"):
            r = subprocess.run([exe, path, fasta], capture_output=True, text=True)
            assert r.returncode == 0 and "roundtrip ok" in r.stdout, (path, fasta, r.stdout, r.stderr)


def test_pack_bases_in_ranges_equals_one_call():
    """The C++ host packs a batch on several threads: ranges cut at multiples of 32 bases, each packed into
    packed + lo / 4, exception positions shifted by lo (src/TGSFilter.cpp).  Same bytes as one call."""
    lib = _capi.load()
    rng = np.random.default_rng(11)
    for n, threads in ((100_003, 4), (1 << 16, 7), (999, 3), (64, 2)):
        b = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].copy()
        hits = rng.integers(0, n, max(3, n // 50))
        b[hits] = np.frombuffer(b"Nacgtn", dtype=np.uint8)[rng.integers(0, 6, len(hits))]
        cap = n + 8

        def pack(lo, hi, packed):
            pos = np.zeros(cap, dtype=np.uint64)
            val = np.zeros(cap, dtype=np.uint8)
            ne = C.c_uint64(0)
            rc = lib.tgsf_pack_bases(b.ctypes.data + lo, hi - lo, packed.ctypes.data + lo // 4, pos.ctypes.data,
                                     val.ctypes.data, cap, C.byref(ne))
            assert rc == 0
            return pos[:ne.value] + lo, val[:ne.value]

        whole = np.zeros(n // 4 + 64, dtype=np.uint8)
        wp, wv = pack(0, n, whole)
        parts = np.zeros(n // 4 + 64, dtype=np.uint8)
        chunk = ((n // threads + 31) // 32) * 32
        pp, pv = [], []
        for t in range(threads):
            lo = min(n, t * chunk)
            hi = n if t == threads - 1 else min(n, lo + chunk)
            if hi > lo:
                p_, v_ = pack(lo, hi, parts)
                pp.append(p_)
                pv.append(v_)
        assert np.array_equal(whole, parts)
        assert np.array_equal(wp, np.concatenate(pp)) and np.array_equal(wv, np.concatenate(pv))
