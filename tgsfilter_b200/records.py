"""Host-side record assembly from the coordinates libtgsf_cuda returns.

The GPU never touches names or output text: ``tgsf_collect`` returns (read, start, len, status)
per kept region and the host formats records exactly like the reference worker does
(T.cpp:2011-2053): the first emitted piece of a read keeps the raw name, later ones get
``newSeqName(rawName, passNum)`` which inserts ":N" before the first whitespace (T.cpp:1680-1701).
"""
from __future__ import annotations

from typing import Iterator, List, Tuple

import numpy as np

from . import _capi

_SPACE = b" \t\n\v\f\r"  # std::isspace in the "C" locale


def new_seq_name(raw_name: bytes, number: int) -> bytes:
    """newSeqName, T.cpp:1680-1701."""
    add = b":%d" % number
    for i, ch in enumerate(raw_name):
        if ch in _SPACE:
            return raw_name[:i] + add + raw_name[i:]
    return raw_name + add


def iter_emitted(pieces: np.ndarray) -> Iterator[Tuple[int, int, int, int]]:
    """Yields (read, start, len, pass_num) for every piece that becomes a record, in order."""
    last_read, pass_num = -1, 1
    for i in range(len(pieces)):
        p = pieces[i]
        r = int(p["read"])
        if r != last_read:
            last_read, pass_num = r, 1
        if int(p["status"]) != _capi.PIECE_EMIT:
            continue
        yield r, int(p["start"]), int(p["len"]), pass_num
        pass_num += 1


def format_records(batch, pieces: np.ndarray, fastq: bool = True) -> List[Tuple[bytes, bytes, int]]:
    """[(record text, record name, seq length)] as enqueued at T.cpp:2033/2050."""
    out = []
    for r, start, ln, pass_num in iter_emitted(pieces):
        name = batch.name(r)
        if pass_num >= 2:
            name = new_seq_name(name, pass_num)
        b, q = batch.read(r)
        seq = b[start:start + ln].tobytes()
        if fastq:
            qs = q[start:start + ln].tobytes() if q is not None else b""
            text = b"@" + name + b"\n" + seq + b"\n+\n" + qs + b"\n"
        else:
            text = b">" + name + b"\n" + seq + b"\n"
        out.append((text, name, ln))
    return out
