"""Readers for yaha's on-disk formats (kept bit-compatible) and for
FASTA/FASTQ read files.  Used by the Python binding, the tests and bench.py.

  .nib2  : Compress.c:28-63 (layout), Compress.c:76-134 (reader), BaseSeq.c:115-125
  index  : Index.c:161-194 (writer), Query.c:599-626 (reader)
"""
from __future__ import annotations

import dataclasses
import numpy as np

NIB2_MARKER = 0x01020304
INDEX_VERSION = 0xFFFFFFFF

# 4-bit codes, Math.c:141-156: T C A G N B D H K M R S V W X Y = 0..15, unknown -> 14 (X)
_CODE_CHARS = b"TCAGNBDHKMRSVWXY"
CODE_OF_CHAR = np.full(256, 14, dtype=np.uint8)
for _i, _c in enumerate(_CODE_CHARS):
    CODE_OF_CHAR[_c] = _i
    CODE_OF_CHAR[ord(chr(_c).lower())] = _i
CODE_OF_CHAR[ord("U")] = 0
CODE_OF_CHAR[ord("u")] = 0
CODE_OF_CHAR[128:] = 14
CHAR_OF_CODE = np.frombuffer(_CODE_CHARS, dtype=np.uint8)
_COMP_CHARS = b"AGTCNVHDMKYSBWXR"            # complement base of code i (Math.c:156)
COMP_CODE = CODE_OF_CHAR[np.frombuffer(_COMP_CHARS, dtype=np.uint8)]


@dataclasses.dataclass
class Nib2:
    names: list[str]
    starts: np.ndarray        # global base offset of each sequence (= 2 * byte offset)
    lengths: np.ndarray
    bases: np.ndarray         # uint8 view of the packed base area (high nibble = even offset)
    max_roff: int

    def base(self, off: int) -> int:
        b = int(self.bases[off >> 1])
        return (b & 15) if (off & 1) else (b >> 4)

    def unpack(self, start: int, n: int) -> np.ndarray:
        """Codes of bases [start, start+n) as one byte each."""
        lo = start >> 1
        hi = (start + n + 1) >> 1
        blk = self.bases[lo:hi]
        both = np.empty(2 * len(blk), dtype=np.uint8)
        both[0::2] = blk >> 4
        both[1::2] = blk & 15
        o = start & 1
        return both[o:o + n]


def load_nib2(path: str) -> Nib2:
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    head = raw[:16].view(np.uint32)
    ver = int(head[1])
    if int(head[0]) != NIB2_MARKER or ver not in (1, 2):
        raise ValueError("Input nib2 file bad header format.")
    bases_off, nseq = int(head[2]), int(head[3])
    blk = 12 if ver == 1 else 16
    rec = raw[16:16 + blk * nseq].view(np.uint32).reshape(nseq, blk // 4)
    name_start = 16 + blk * nseq + 4
    names, starts, lengths = [], [], []
    for i in range(nseq):
        if ver == 1:
            info = int(rec[i, 2])
            noff, nlen = (info >> 16) & 0xFFFF, info & 0xFFFF
        else:
            noff, nlen = int(rec[i, 2]), int(rec[i, 3])
        names.append(bytes(raw[name_start + noff:name_start + noff + nlen]).decode())
        starts.append(2 * int(rec[i, 0]))
        lengths.append(int(rec[i, 1]))
    return Nib2(names, np.array(starts, dtype=np.int64), np.array(lengths, dtype=np.int64),
                raw[bases_off:], starts[-1] + lengths[-1])


@dataclasses.dataclass
class Index:
    word_len: int
    max_hits: int
    total: int
    so: np.ndarray            # uint32[4^K + 1]
    roa: np.ndarray           # uint32[total]


def load_index(path: str) -> Index:
    raw = np.memmap(path, dtype=np.uint32, mode="r")
    if int(raw[0]) != INDEX_VERSION:
        raise ValueError("Index file version is out of date.")
    k, mh, tot = int(raw[1]), int(raw[2]), int(raw[3])
    nso = 4 ** k + 1
    return Index(k, mh, tot, raw[4:4 + nso], raw[4 + nso:])


def encode(seq: bytes | np.ndarray) -> np.ndarray:
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else seq
    return CODE_OF_CHAR[a]


def revcomp_codes(fwd: np.ndarray) -> np.ndarray:
    return COMP_CODE[fwd[::-1]]


def read_queries(path: str, word_len: int = 15, max_len: int = 32000):
    """FASTA/FASTQ reader with the reference's skipping rules (Query.c:102-228):
    names cut at newline with spaces -> '_', reads longer than max_len or shorter than
    word_len are skipped.  Returns list of (name, bases_bytes)."""
    data = open(path, "rb").read()
    out = []
    if not data:
        return out
    fastq = data[:1] == b"@"
    if fastq:
        lines = data.split(b"\n")
        i = 0
        while i < len(lines):
            if not lines[i].startswith(b"@"):
                i += 1
                continue
            name = lines[i][1:]
            i += 1
            seq = []
            while i < len(lines) and not lines[i].startswith(b"+"):
                seq.append(lines[i])
                i += 1
            s = b"".join(seq)
            i += 1
            q = 0
            while i < len(lines) and q < len(s):
                q += len(lines[i])
                i += 1
            if word_len <= len(s) <= max_len:
                out.append((name.replace(b" ", b"_").decode(), s))
    else:
        for rec in data.split(b">")[1:]:
            nl = rec.find(b"\n")
            name = rec[:nl] if nl >= 0 else rec
            s = rec[nl + 1:].replace(b"\n", b"") if nl >= 0 else b""
            if word_len <= len(s) <= max_len:
                out.append((name.replace(b" ", b"_").decode(), s))
    return out
