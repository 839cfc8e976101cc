import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.getcwd())
from tgsfilter_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ext = sys.argv[2] if len(sys.argv) > 2 else ".fq"   # ".fq.gz": per-record gzip members like the reference
in_gz = len(sys.argv) > 3 and sys.argv[3] in ("gzin", "bgzfin")  # gzip -6 input: 16 plain members, or BGZF blocks
in_bgzf = len(sys.argv) > 3 and sys.argv[3] == "bgzfin"


def bgzf_compress(data, blk=65280):
    import struct, zlib
    out = bytearray()
    for i in range(0, len(data), blk):
        chunk = data[i:i + blk]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        d = c.compress(chunk) + c.flush()
        out += (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(d) + 25) + d +
                struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return bytes(out)


batch = synth.make_config(2, n, with_names=False)
d = "/dev/shm/f2f"; os.makedirs(d, exist_ok=True)
fq = os.path.join(d, "in.fq")
with open(fq, "wb") as f: f.write(batch.to_fastq())
n_bases = batch.n_bases
del batch
if in_gz:
    import gzip
    from concurrent.futures import ProcessPoolExecutor
    raw = open(fq, "rb").read()
    # cut at record starts ("\n@read" only occurs at a header in this synthetic file: qualities are < '@')
    cuts = [0]
    for k in range(1, 16):
        cuts.append(raw.index(b"\n@read", k * len(raw) // 16) + 1)
    cuts.append(len(raw))
    with ProcessPoolExecutor(16) as ex:
        if in_bgzf:
            parts = list(ex.map(bgzf_compress, [raw[a:b] for a, b in zip(cuts[:-1], cuts[1:])]))
        else:
            parts = list(ex.map(gzip.compress, [raw[a:b] for a, b in zip(cuts[:-1], cuts[1:])], [6] * 16))
    fq = fq + ".gz"
    with open(fq, "wb") as f:
        for p_ in parts:
            f.write(p_)
    del raw, parts
print("bases", n_bases, "file MB", os.path.getsize(fq) / 1e6)
if "prepare_only" in sys.argv:
    sys.exit(0)
for name, exe, extra in (("host_b200", "src/tgsfilter", []), ("reference", "oracle/_ref/tgsfilter", ["-t", str(min(32, (os.cpu_count() or 2) - 1))])):
    out = os.path.join(d, name + ext)
    for rep in range(2):
        t0 = time.perf_counter()
        pr = subprocess.run([exe, "-i", fq, "-x", "ont", "-o", out] + extra, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        print(name, "rc", pr.returncode, "%.2f s" % dt, "%.3f Gbases/s" % (n_bases / dt / 1e9))
def slurp(path):
    import gzip
    if path.endswith(".gz"):  # streaming reader: gzip.decompress() is quadratic in the number of members
        with gzip.open(path, "rb") as f:
            return f.read()
    return open(path, "rb").read()
a = slurp(os.path.join(d, "host_b200" + ext)); b = slurp(os.path.join(d, "reference" + ext))
print("same multiset of records:", sorted(a.split(b"@read")) == sorted(b.split(b"@read")), len(a), len(b))
