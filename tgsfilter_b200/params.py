"""Host-side mirror of the reference's parameter block for the per-read path.

``FilterParams`` carries the subset of ``Para_A24`` (T.cpp:82-172) that ``filter_sequence`` /
``adapterMap`` / ``GetEditDistance`` read, with the constructor defaults of T.cpp:129-171, the
per-read-type similarity defaults of T.cpp:438-457 and the clamps of ``TGSFilter_cmd``
(T.cpp:228-331).  ``ADAPTER_LIB`` is the built-in library of T.cpp:2969-2991; ``rev_comp``
follows ``rev_comp_seq`` / the ``complement`` table (T.cpp:859-867, 2954-2967).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Sequence

import numpy as np

from . import _capi

ADAPTER_LIB: List[bytes] = [
    b"ATCTCTCTCTTTTCCTCCTCCTCCGTTGTTGTTGTTGAGAGAGAT",  # PacBio blunt
    b"ATCTCTCTCAACAACAACAACGGAGGAGGAGGAAAAGAGAGAGAT",
    b"AAAAAAAAAAAAAAAAAATTAACGGAGGAGGAGGA",  # PacBio C2 primer
    b"TCCTCCTCCTCCGTTAATTTTTTTTTTTTTTTTTT",
    b"AATGTACTTCGTTCAGTTACGTATTGCT",  # ONT ligation
    b"AGCAATACGTAACTGAACGAAGTACATT",
    b"GCAATACGTAACTGAACGAAGT",
    b"ACTTCGTTCAGTTACGTATTGC",
    b"GTTTTCGCATTTATCGTGAAACGCTTTCGCGTTTTTCGTGCGCCGCTTCA",  # ONT rapid
    b"TGAAGCGGCGCACGAAAAACGCGAAAGCGTTTCACGATAAATGCGAAAAC",
    b"GGCGTCTGCTTGGGTGTTTAACCTTTTTGTCAGAGAGGTTCCAAGTCAGAGAGGTTCCT",  # 1D^2
    b"AGGAACCTCTCTGACTTGGAACCTCTCTGACAAAAAGGTTAAACACCCAAGCAGACGCC",
    b"GGAACCTCTCTGACTTGGAACCTCTCTGACAAAAAGGTTAAACACCCAAGCAGACGCCAGCAAT",
    b"ATTGCTGGCGTCTGCTTGGGTGTTTAACCTTTTTGTCAGAGAGGTTCCAAGTCAGAGAGGTTCC",
    b"TTTTTTTTCCTGTACTTCGTTCAGTTACGTATTGCT",  # LA / NA / RA / RAT top strand
    b"AGCAATACGTAACTGAACGAAGTACAGGAAAAAAAA",
    b"GCAATACGTAACTGAACGAAGTACAGG",  # LA bottom strand
    b"CCTGTACTTCGTTCAGTTACGTATTGC",
    b"ACGTAACTGAACGAAGTACAGG",  # NA bottom strand
    b"CCTGTACTTCGTTCAGTTACGT",
    b"CTTGCGGGCGGCGGACTCTCCTCTGAAGATAGAGCGACAGGCAAG",  # cDNA RT adapter
    b"CTTGCCTGTCGCTCTATCTTCAGAGGAGAGTCCGCCGCCCGCAAG",
]

_COMPLEMENT = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in ("AT", "GC", "CG", "TA", "at", "gc", "cg", "ta", "MK", "RY", "WW", "SS", "YR", "KM",
               "mk", "ry", "ww", "ss", "yr", "km"):
    _COMPLEMENT[ord(_a)] = ord(_b)


def rev_comp(seq: bytes) -> bytes:
    """rev_comp_seq (T.cpp:860-867): unknown bytes become 'N'."""
    arr = np.frombuffer(seq, dtype=np.uint8)
    return _COMPLEMENT[arr[::-1]].tobytes()


def rev_comp_rows(rows: np.ndarray) -> np.ndarray:
    """Reverse-complement every row of a (n, L) uint8 matrix."""
    return _COMPLEMENT[rows[:, ::-1]]


@dataclasses.dataclass
class FilterParams:
    min_len: int = 1000            # -l
    max_len: int = 2147483647      # -L
    min_q: float = -1.0            # -q (pre-pass picks the default, T.cpp:1066-1075)
    max_q: float = 255.0           # -Q
    bc_len: int = 150              # -e
    head_trim: int = -1            # -5 (<0: decided by the pre-pass)
    tail_trim: int = -1            # -3
    end_len: int = 150             # -E
    end_match_len: int = 4         # -m (usage text says 15, constructor says 4: T.cpp:55 vs 148)
    mid_match_len: int = 35        # -M
    extra_len: int = 50            # -T
    end_sim: float = 0.0           # -s
    mid_sim: float = 0.0           # -S
    kmer: int = 11                 # -k
    min_repeat: int = 0            # -p
    qtype: int = 33
    filter: bool = True
    only_qc: bool = False
    discard: bool = False          # -D
    gz_blocks: bool = False        # also deflate-encode the emitted pieces on the GPU (tgsf_collect_gz)
    gz_fasta: bool = False         # ... as FASTA records
    adapters: Sequence[bytes] = ()
    max_read_len: int = 0
    n_slots: int = 0

    def apply_read_type(self, read_type: str) -> "FilterParams":
        """Per-type similarity defaults, T.cpp:438-457 (float32 values)."""
        mid = {"hifi": 0.95, "clr": 0.9, "ont": 0.9}
        end = {"hifi": 0.9, "clr": 0.8, "ont": 0.75}
        if self.mid_sim == 0:
            self.mid_sim = mid[read_type]
        if self.end_sim == 0:
            self.end_sim = end[read_type]
        return self

    @property
    def flags(self) -> int:
        return ((_capi.FLAG_FILTER if self.filter else 0)
                | (_capi.FLAG_ONLY_QC if self.only_qc else 0)
                | (_capi.FLAG_DISCARD_MID if self.discard else 0)
                | (_capi.FLAG_GZ_BLOCKS if self.gz_blocks else 0)
                | (_capi.FLAG_GZ_FASTA if self.gz_fasta else 0))

    def to_c(self):
        """Returns (tgsf_params struct, keep-alive objects)."""
        n = len(self.adapters)
        seqs = (C.c_char_p * max(n, 1))(*[bytes(a) for a in self.adapters])
        lens = (C.c_int32 * max(n, 1))(*[len(a) for a in self.adapters])
        p = _capi.Params(
            min_len=self.min_len, max_len=self.max_len, min_q=self.min_q, max_q=self.max_q,
            bc_len=self.bc_len, head_trim=self.head_trim, tail_trim=self.tail_trim,
            end_len=self.end_len, end_match_len=self.end_match_len,
            mid_match_len=self.mid_match_len, extra_len=self.extra_len, end_sim=self.end_sim,
            mid_sim=self.mid_sim, kmer=self.kmer, min_repeat=self.min_repeat, qtype=self.qtype,
            flags=self.flags, n_adapters=n,
            adapter_seq=C.cast(seqs, C.POINTER(C.c_char_p)),
            adapter_len=C.cast(lens, C.POINTER(C.c_int32)),
            max_read_len=self.max_read_len, n_slots=self.n_slots)
        return p, (seqs, lens)
