"""Access to the UNMODIFIED reference built into oracle/_ref/ (test infrastructure only).

``oracle/_ref/ref_driver`` exposes the reference's edlibAlign, TGSFilterTask::filter_sequence and
pre-pass bodies on binary in/out files (see oracle/ref_driver.cpp for the format);
``oracle/_ref/tgsfilter`` is the reference CLI.  Both exist only where /root/reference was
available at build time, or where the prebuilt files travelled (gpurun snapshot).
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
DRIVER = os.path.join(REF_DIR, "ref_driver")
CLI = os.path.join(REF_DIR, "tgsfilter")


def available() -> bool:
    if not (os.path.exists(DRIVER) and os.path.exists(CLI)) and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(DRIVER) and os.path.exists(CLI)


def _pstr(b: bytes) -> bytes:
    return struct.pack("<I", len(b)) + b


class _Rd:
    def __init__(self, data: bytes):
        self.d, self.p = data, 0

    def get(self, fmt):
        v = struct.unpack_from("<" + fmt, self.d, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v

    def arr(self, dtype, n):
        a = np.frombuffer(self.d, dtype=dtype, count=n, offset=self.p)
        self.p += a.nbytes
        return a

    def str(self):
        n = self.get("I")
        s = self.d[self.p:self.p + n]
        self.p += n
        return s


def _run_driver(mode: str, payload: bytes) -> bytes:
    with tempfile.TemporaryDirectory() as td:
        fi, fo = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fi, "wb") as f:
            f.write(payload)
        subprocess.run([DRIVER, mode, fi, fo], check=True, stderr=subprocess.DEVNULL)
        with open(fo, "rb") as f:
            return f.read()


def edlib_batch(cases):
    """cases: [(query, target, k)] -> [(d, aln_len, [(start, end), ...])] from the real edlibAlign."""
    payload = [struct.pack("<I", len(cases))]
    for q, t, k in cases:
        payload.append(struct.pack("<i", k) + _pstr(q) + _pstr(t))
    r = _Rd(_run_driver("edlib", b"".join(payload)))
    out = []
    for _ in cases:
        d, n, alen = r.get("iii")
        locs = r.arr("<i4", 2 * n).reshape(n, 2)
        out.append((d, alen, [tuple(x) for x in locs.tolist()]))
    return out


def _params_blob(p, outfq: int) -> bytes:
    blob = struct.pack("<iiffiiiiiiiffiiiIi", p.min_len, p.max_len, p.min_q, p.max_q, p.bc_len,
                       p.head_trim, p.tail_trim, p.end_len, p.end_match_len, p.mid_match_len,
                       p.extra_len, p.end_sim, p.mid_sim, p.kmer, p.min_repeat, p.qtype, p.flags,
                       outfq)
    blob += struct.pack("<i", len(p.adapters))
    for a in p.adapters:
        blob += _pstr(bytes(a))
    return blob


TABLES = ("raw5p_cnt", "raw5p_qual", "raw3p_cnt", "raw3p_qual", "clean5p_cnt", "clean5p_qual",
          "clean3p_cnt", "clean3p_qual", "raw_bin_cnt", "raw_bin_qual", "clean_bin_cnt",
          "clean_bin_qual")


def perread(params, batch, outfq: int = 1):
    """Run the reference worker body over a batch.  Returns dict with drop_info, raw_hist,
    clean_hist, the 12 tables (rows x 5 uint64) and records [(text, name, seqlen)]."""
    payload = [_params_blob(params, outfq), struct.pack("<I", batch.n_reads)]
    for i in range(batch.n_reads):
        b, q = batch.read(i)
        payload.append(_pstr(batch.name(i)) + _pstr(b.tobytes())
                       + _pstr(q.tobytes() if q is not None else b""))
    r = _Rd(_run_driver("perread", b"".join(payload)))
    out = {"drop_info": r.arr("<u8", 17).copy(), "raw_hist": r.arr("<u8", 256).copy(),
           "clean_hist": r.arr("<u8", 256).copy()}
    for name in TABLES:
        rows = r.get("I")
        out[name] = r.arr("<u8", rows * 5).reshape(rows, 5).copy()
    nrec = r.get("I")
    recs = []
    for _ in range(nrec):
        text = r.str()
        name = r.str()
        slen = r.get("i")
        recs.append((text, name, slen))
    out["records"] = recs
    return out


def prepass(end_len, bc_len, end_bias, mid_sim, ends5p: np.ndarray, ends3p: np.ndarray, lib):
    n, row = ends5p.shape
    payload = [struct.pack("<iiffII", end_len, bc_len, end_bias, mid_sim, n, row),
               struct.pack("<i", len(lib))]
    payload += [_pstr(a) for a in lib]
    payload += [_pstr(ends5p[i].tobytes()) for i in range(n)]
    payload += [_pstr(ends3p[i].tobytes()) for i in range(n)]
    r = _Rd(_run_driver("prepass", b"".join(payload)))
    t5, t3 = r.get("ii")
    a5, a3 = r.str(), r.str()
    d5, d3 = r.get("ff")
    return {"trim5p": t5, "trim3p": t3, "adapter5p": a5, "adapter3p": a3, "dep5p": d5, "dep3p": d3}


def run_cli(args, input_bytes: bytes, in_name: str = "in.fq", out_name: str | None = "out.fq",
            timeout: int = 600):
    """Run the reference CLI on an input file.  Returns (returncode, output bytes, stderr text,
    html text)."""
    with tempfile.TemporaryDirectory() as td:
        fi = os.path.join(td, in_name)
        with open(fi, "wb") as f:
            f.write(input_bytes)
        cmd = [CLI, "-i", fi] + list(args)
        fo = None
        if out_name is not None:
            fo = os.path.join(td, out_name)
            cmd += ["-o", fo]
        pr = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout,
                            cwd=td)
        out = pr.stdout
        if fo is not None and os.path.exists(fo):
            with open(fo, "rb") as f:
                out = f.read()
        html = ""
        for fn in os.listdir(td):
            if fn.endswith(".html"):
                with open(os.path.join(td, fn), "r", errors="replace") as f:
                    html = f.read()
        return pr.returncode, out, pr.stderr.decode("utf-8", "replace"), html
