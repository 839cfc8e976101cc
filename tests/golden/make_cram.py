"""(Re)generate tests/golden/reads_cram.bam: 48 synthetic ONT reads as CRAM in a file NAMED .bam — the only way CRAM
reaches the reference, whose GetFileType knows the suffixes sam / bam only (T.cpp:839-857) while hts_open sniffs the
content.  Needs an htslib prefix (HTS_DIR, default /root/reference with its vendored lib/libhts.a); run from the repo
root on the build container.  The reads are synth.make_config(2, 48, max_len=9000, with_names=False)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bam_lib  # noqa: E402
from tgsfilter_b200 import synth  # noqa: E402

HTS = os.environ.get("HTS_DIR", "/root/reference")
with tempfile.TemporaryDirectory() as td:
    exe = os.path.join(td, "mkcram")
    subprocess.run(["g++", "-O2", "-no-pie", "-w", "-I" + HTS + "/include", os.path.join(ROOT, "tests", "cpp", "mkcram.cpp"),
                    HTS + "/lib/libhts.a", HTS + "/lib/libdeflate.a", HTS + "/lib/libisal.a", "-lz", "-pthread", "-o", exe],
                   check=True)
    fq = synth.make_config(2, 48, max_len=9000, with_names=False).to_fastq()
    bam, _ = bam_lib.from_fastq(fq)
    src = os.path.join(td, "in.bam")
    with open(src, "wb") as f:
        f.write(bam)
    out = os.path.join(ROOT, "tests", "golden", "reads_cram.bam")
    subprocess.run([exe, src, out], check=True)
    print(out, os.path.getsize(out), "bytes")
