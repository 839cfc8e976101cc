"""Single-stream gzip input (src/pinflate.hpp): decoder alone, ingest to parsed batches, and the CLI file to file.

usage: python profiles/pinflate_bench.py [MB of FASTQ, default 200] [cli]
Builds two one-member .fastq.gz files in /dev/shm: `gzip -1` (matches cross the whole stream) and a pigz-style
stream (128 KB chunks deflated independently, each closed by a sync flush).  Needs build/ingest_check and
build/pinflate_check (g++ -O2 -std=c++17 -pthread tests/cpp/<name>.cpp -lz).  With `cli`: also runs
src/tgsfilter (needs a GPU) with the parallel and the sequential decoder and, if present, the reference CLI.
"""
import concurrent.futures as cf
import os
import struct
import subprocess
import sys
import time
import zlib

import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
work = "/dev/shm/pinflate_bench"
os.makedirs(work, exist_ok=True)
rng = np.random.default_rng(1)
t0 = time.time()
parts, tot, i = [], 0, 0
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
while tot < mb * 1_000_000:
    ln = int(rng.lognormal(9.5, 0.8)) + 200
    seq = acgt[rng.integers(0, 4, ln)].tobytes()
    q = (rng.normal(20, 6, ln).clip(2, 40).astype(np.uint8) + 33).tobytes()
    parts.append(b"@read%d runid=abcdef ch=%d\n" % (i, i % 512) + seq + b"\n+\n" + q + b"\n")
    tot += 2 * ln
    i += 1
fq = b"".join(parts)
open(work + "/in.fq", "wb").write(fq)
t1 = time.time()
subprocess.run("gzip -1 -c %s/in.fq > %s/gzip1.fq.gz" % (work, work), shell=True, check=True)
t2 = time.time()


def chunk(args):
    data, last = args
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    return c.compress(data) + (c.flush(zlib.Z_FINISH) if last else c.flush(zlib.Z_SYNC_FLUSH))


cuts = list(range(0, len(fq), 131072))
with cf.ThreadPoolExecutor(os.cpu_count()) as ex:
    blobs = list(ex.map(chunk, [(fq[a:a + 131072], a == cuts[-1]) for a in cuts]))
with open(work + "/pigz.fq.gz", "wb") as f:
    f.write(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + b"".join(blobs) + struct.pack("<II", zlib.crc32(fq), len(fq) & 0xFFFFFFFF))
t3 = time.time()
print("FASTQ %d bytes, %d reads (%.1f s); gzip -1: %d bytes (%.1f s); pigz-style: %d bytes (%.1f s); %d cores" % (
    len(fq), i, t1 - t0, os.path.getsize(work + "/gzip1.fq.gz"), t2 - t1, os.path.getsize(work + "/pigz.fq.gz"), t3 - t2,
    os.cpu_count()), flush=True)


def timed(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    a = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, env=e, timeout=120)
    return time.time() - a, r


for name in ("gzip1", "pigz"):
    path = "%s/%s.fq.gz" % (work, name)
    for thr in (4, 8, max(1, (os.cpu_count() or 2) - 2)):
        dt, r = timed([root + "/build/pinflate_check", path, str(thr)])
        print("%-6s decoder alone, %2d workers: rc %d  %s | %s" % (name, thr, r.returncode, r.stdout.strip(), r.stderr.strip()), flush=True)
    for thr in (1, 2, 4, 8, 12):
        dt, r = timed([root + "/build/ingest_check", path, "1", "67108864", "1"],
                      {"TGSF_INFLATE_THREADS": str(thr), "INGEST_ONLY": "serial"})
        print("%-6s ingest to batches, TGSF_INFLATE_THREADS=%2d: %.2f s  (%s)" % (name, thr, dt, " ".join(r.stdout.split()[:2])), flush=True)
    dt, r = timed([root + "/build/ingest_check", path, "1", "67108864", "1"], {"TGSF_ZLIB_INFLATE": "1", "INGEST_ONLY": "serial"})
    print("%-6s ingest to batches, zlib: %.2f s" % (name, dt), flush=True)

if "cli" in sys.argv[2:]:
    path = work + "/gzip1.fq.gz"
    outs = []
    for label, env in (("parallel inflate (default)", {}), ("sequential inflate", {"TGSF_SERIAL_INFLATE": "1"}),
                       ("parallel inflate (default)", {})):
        out = "%s/out_%d.fq" % (work, len(outs))
        dt, r = timed([root + "/src/tgsfilter", "-i", path, "-o", out, "-x", "ont"], dict(env, TGSF_TIMING="1"))
        outs.append(out)
        print("CLI .gz -> plain, %s: %.2f s rc %d | %s" % (label, dt, r.returncode,
              " ; ".join(l for l in r.stderr.splitlines() if "TIMING" in l.upper() or "time" in l.lower())[:600]), flush=True)
    same = open(outs[0], "rb").read() == open(outs[1], "rb").read()
    print("outputs identical:", same, os.path.getsize(outs[0]), flush=True)
    ref = root + "/oracle/_ref/tgsfilter"
    if os.path.exists(ref):
        dt, r = timed([ref, "-i", path, "-o", work + "/ref.fq", "-x", "ont", "-t", str(max(1, (os.cpu_count() or 2) - 1))])
        print("reference CLI .gz -> plain: %.2f s rc %d" % (dt, r.returncode), flush=True)
