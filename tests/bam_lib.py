"""Minimal BAM / SAM writers for the ingest tests (SAM specification v1.6, sections 1.4 and 4.2): unaligned
records in BGZF blocks.  tests/test_host.py checks the files with the compiled reference (htslib) where it is
available: the reference CLI must treat them exactly like the equivalent FASTQ."""
import struct
import zlib

_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_BASE = [0, 65, 67, 0, 71, 0, 0, 0, 84, 0, 0, 0, 0, 0, 0, 78]  # the reference's nibble -> byte table (T.cpp:31)


def through_reference_table(seq: str) -> bytes:
    """What read_bam (T.cpp:1896-1903) makes of these SEQ characters."""
    return bytes(_BASE[_CODE.get(c.upper(), 15)] for c in seq)


def bgzf_block(data: bytes) -> bytes:
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = 12 + 6 + len(comp) + 8 - 1
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + comp +
            struct.pack("<II", zlib.crc32(data), len(data)))


def bgzf(data: bytes, blk: int = 0xff00) -> bytes:
    return b"".join(bgzf_block(data[i:i + blk]) for i in range(0, len(data), blk)) + bgzf_block(b"")


def bam_record(name: str, seq: str, qual, flag: int = 4, tags: bytes = b"") -> bytes:
    n = len(seq)
    packed = bytearray((n + 1) // 2)
    for i, ch in enumerate(seq):
        v = _CODE.get(ch.upper(), 15)
        packed[i >> 1] |= (v << 4) if i % 2 == 0 else v
    q = bytes(qual) if qual is not None else b"\xff" * n
    nm = name.encode() + b"\0"
    core = struct.pack("<iiBBHHHiiii", -1, -1, len(nm), 0, 4680, 0, flag, n, -1, -1, 0)
    body = core + nm + bytes(packed) + q + tags
    return struct.pack("<i", len(body)) + body


def bam_file(records, text: bytes = b"@HD\tVN:1.6\tSO:unknown\n", refs=()) -> bytes:
    h = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, length in refs:
        h += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", length)
    return bgzf(h + b"".join(records))


def sam_line(name: str, seq: str, qual, flag: int = 4, tags: str = "") -> bytes:
    q = b"*" if qual is None or len(seq) == 0 else bytes(int(v) + 33 for v in qual)
    return ("%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t" % (name, flag, seq if seq else "*")).encode() + q + \
        (("\t" + tags).encode() if tags else b"") + b"\n"


def from_fastq(fq: bytes):
    """(bam, sam) holding the records of a 4-line FASTQ (Phred+33)."""
    lines = fq.split(b"\n")
    recs, sam = [], [b"@HD\tVN:1.6\tSO:unknown\n"]
    for i in range(0, len(lines) - 3, 4):
        name = lines[i][1:].split()[0].decode()
        seq = lines[i + 1].decode()
        qual = [c - 33 for c in lines[i + 3]]
        recs.append(bam_record(name, seq, qual, flag=(4, 0, 16, 256)[(i // 4) % 4], tags=b"RGZgrp\0" if i % 3 else b""))
        sam.append(sam_line(name, seq, qual, tags="RG:Z:grp" if i % 3 else ""))
    return bam_file(recs, refs=(("chr1", 1000),)), b"".join(sam)
