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


# ---------------------------------------------------------------------------------------------
# Writers (small inputs; the tests use them so that no reference binary is needed at run time)
# ---------------------------------------------------------------------------------------------
def read_fasta(path: str):
    """[(name, bases_bytes)]; names are cut at the first space (Compress.c:277-284)."""
    out = []
    for rec in open(path, "rb").read().split(b">")[1:]:
        nl = rec.find(b"\n")
        name = rec[:nl].split(b" ")[0]
        out.append((name.decode(), rec[nl + 1:].replace(b"\n", b"").replace(b"\r", b"")))
    return out


def build_nib2(seqs) -> bytes:
    """.nib2 v2 image for [(name, bases)] (layout: Compress.c:28-63,140-191,199-218)."""
    names = b"".join(n.encode() for n, _ in seqs)
    names_pad = names + b"\0" * (-len(names) % 4)
    nseq = len(seqs)
    bases_off = 20 + 16 * nseq + len(names_pad)
    recs, blobs = [], []
    byte_off, name_off = 0, 0
    for n, s in seqs:
        codes = encode(s)
        L = len(codes)
        pad = (-L) % 8                      # pad with X (14) to a 4-byte boundary (Compress.c:206-215)
        c = np.concatenate([codes, np.full(pad, 14, np.uint8)])
        packed = (c[0::2] << 4) | c[1::2]
        recs.append(np.array([byte_off, L, name_off, len(n.encode())], dtype="<u4").tobytes())
        blobs.append(packed.astype(np.uint8).tobytes())
        byte_off += len(packed)
        name_off += len(n.encode())
    head = np.array([NIB2_MARKER, 2, bases_off, nseq], dtype="<u4").tobytes()
    return head + b"".join(recs) + np.array([0], dtype="<u4").tobytes() + names_pad + b"".join(blobs)


def _marsaglia(state):
    """Math.c:274-284 xorshift step; state is a list of 5 python ints (uint32)."""
    M = 0xFFFFFFFF
    t = state[0] ^ (state[0] >> 7)
    state[0], state[1], state[2], state[3] = state[1], state[2], state[3], state[4]
    state[4] = ((state[4] ^ ((state[4] << 6) & M)) ^ (t ^ ((t << 13) & M))) & M
    return ((state[1] + state[1] + 1) * state[4]) & M


def _rand_sample(state, inp: np.ndarray, out_len: int) -> np.ndarray:
    """Order-preserving Floyd sample (Math.c:304-343)."""
    n = len(inp)
    marked = np.zeros(n, dtype=bool)
    keep_marked, select = True, out_len
    if out_len > n // 2:
        keep_marked, select = False, n - out_len
    for i in range(n - select, n):
        pos = int((_marsaglia(state) / 4294967296.0) * (i + 1))
        if marked[pos]:
            marked[i] = True
        else:
            marked[pos] = True
    return inp[marked == keep_marked]


def _indexed_positions(codes: np.ndarray, st: int, ln: int, K: int, skip: int) -> np.ndarray:
    """Global offsets of the k-mers the reference indexes in one sequence (Index.c:95-128): from the sequence start
    every `skip`-th window while it is free of non-ACGT codes; a window that holds one makes the walk resume at the
    first multiple of `skip` (of the GLOBAL offset) at or behind the first good base after that run of bad codes.
    `codes` is the whole genome's code array, one byte per base."""
    end = st + ln - K                                   # last window start of the sequence
    n = len(codes)
    bad = codes > 3
    idx = np.arange(n, dtype=np.int64)
    big = np.int64(1) << 40
    next_bad = np.minimum.accumulate(np.where(bad, idx, big)[::-1])[::-1]          # first bad code at or behind i
    next_good = np.minimum.accumulate(np.where(~bad, idx, big)[::-1])[::-1]
    out = []
    b = st
    while b <= end:
        nb = int(next_bad[b]) if b < n else int(big)
        last = min(end, nb - K)                         # windows [p, p+K) with p <= last are clean
        if last >= b:
            cnt = (last - b) // skip + 1
            out.append(b + skip * np.arange(cnt, dtype=np.int64))
            b += skip * cnt
            if b > end or nb >= big:
                break
            if b + K <= nb:                             # (cannot happen: b is the first window start past `last`)
                continue
        # the window at b holds the bad code at nb: skip the run of bad codes, renormalise (Index.c:110-116)
        x = nb + 1
        x = int(next_good[x]) if x < n else int(big)
        if x >= big:
            break
        b = ((x + skip - 1) // skip) * skip
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def build_index(nib: Nib2, word_len: int, max_hits: int = 65525, skip: int = 1) -> bytes:
    """Index image (Index.c:95-331) for a loaded .nib2, any -L / -S / -H."""
    K = word_len
    total = int(nib.starts[-1] + nib.lengths[-1]) if len(nib.starts) else 0
    n_codes = min(2 * len(nib.bases), total + 16)
    allc = nib.unpack(0, n_codes)
    pos_all, hash_all = [], []
    for st, ln in zip(nib.starts, nib.lengths):
        st, ln = int(st), int(ln)
        if ln < K:
            continue
        p = _indexed_positions(allc, st, ln, K, skip)
        h = np.zeros(len(p), dtype=np.uint32)
        for k in range(K):
            h = (h << 2) | (allc[p + k].astype(np.uint32) & 3)
        pos_all.append(p.astype(np.uint32))
        hash_all.append(h)
    pos = np.concatenate(pos_all) if pos_all else np.zeros(0, np.uint32)
    hsh = np.concatenate(hash_all) if hash_all else np.zeros(0, np.uint32)
    order = np.argsort(hsh, kind="stable")              # ascending offset inside each k-mer list
    roa = pos[order]
    counts = np.bincount(hsh, minlength=4 ** K).astype(np.int64)
    so = np.zeros(4 ** K + 1, dtype=np.int64)
    np.cumsum(counts, out=so[1:])
    over = np.nonzero(counts > max_hits)[0]
    if len(over):                                       # pass 3: down-sample (Index.c:271-315)
        state = [123456789, 362436069, 521288629, 88675123, 886756453]
        pieces, new_counts = [], counts.copy()
        prev = 0
        for hcode in over:
            a, b = int(so[hcode]), int(so[hcode + 1])
            pieces.append(roa[prev:a])
            pieces.append(_rand_sample(state, roa[a:b], max_hits))
            new_counts[hcode] = max_hits
            prev = b
        pieces.append(roa[prev:])
        roa = np.concatenate(pieces)
        np.cumsum(new_counts, out=so[1:])
    head = np.array([INDEX_VERSION, K, max_hits, len(roa)], dtype="<u4")
    return head.tobytes() + so.astype("<u4").tobytes() + roa.astype("<u4").tobytes()


def index_file_name(stem: str, word_len: int, skip: int, max_hits: int) -> str:
    """Main.c:559-563: <stem>.X<LL>_<SS>_<HHHHH>S"""
    return f"{stem}.X{word_len:02d}_{skip:02d}_{max_hits:05d}S"


def write_index(path: str, idx: Index) -> None:
    """Write an Index in the reference's file layout (Index.c:171-176,331)."""
    with open(path, "wb") as f:
        np.array([INDEX_VERSION, idx.word_len, idx.max_hits, len(idx.roa)], dtype="<u4").tofile(f)
        np.ascontiguousarray(idx.so, dtype="<u4").tofile(f)
        np.ascontiguousarray(idx.roa, dtype="<u4").tofile(f)


def main(argv=None) -> int:
    """`python -m yaha_b200.refio -g genome.(fa|nib2) [-L wordLen] [-S skipDist] [-H maxHits]` -- the reference's index
    creation mode (Main.c:567-634): writes <stem>.nib2 (from FASTA input) and <stem>.X<LL>_<SS>_<HHHHH>S next to the
    input, byte for byte the files `yaha -g` writes (defaults -L 15 -S 1 -H 65525, AlignArgs.c:48-87)."""
    import argparse
    import os
    import sys
    ap = argparse.ArgumentParser(prog="python -m yaha_b200.refio", description="yaha-compatible .nib2 and index files")
    ap.add_argument("-g", required=True, metavar="genome.(fa|nib2)")
    ap.add_argument("-L", type=int, default=15, metavar="wordLen")
    ap.add_argument("-S", type=int, default=1, metavar="skipDist")
    ap.add_argument("-H", type=int, default=65525, metavar="maxHits")
    a = ap.parse_args(argv)
    if not 1 <= a.S <= a.L or not 1 <= a.L <= 15:
        print("wordLen must be 1..15 and skipDist 1..wordLen", file=sys.stderr)
        return 1
    stem, ext = os.path.splitext(a.g)
    nib_path = a.g if ext == ".nib2" else stem + ".nib2"
    if ext != ".nib2":
        print(f"Compressing {a.g} into {nib_path}.", file=sys.stderr)
        with open(nib_path, "wb") as f:
            f.write(build_nib2(read_fasta(a.g)))
    idx_path = index_file_name(stem, a.L, a.S, min(a.H, 65525))
    print(f"Creating index file {idx_path}.", file=sys.stderr)
    with open(idx_path, "wb") as f:
        f.write(build_index(load_nib2(nib_path), a.L, max_hits=min(a.H, 65525), skip=a.S))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
