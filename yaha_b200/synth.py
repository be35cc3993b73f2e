"""Seeded synthetic references and reads for the BASELINE.json configs.

Pure numpy; deterministic for a given seed.  The shapes follow SURVEY.md section 8(d):
i.i.d. uniform ACGT references, reads drawn at uniform positions on a random strand with a
per-base error rate split 80 % substitution / 10 % insertion / 10 % deletion.

This module only writes FASTA files; encoding/indexing is done by the host program
(`yaha_b200 -g`, format-compatible with the reference's `yaha -g`, Index.c:49 / Compress.c:220).
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def random_reference(n_bases: int, seed: int) -> np.ndarray:
    """Return `n_bases` i.i.d. uniform ACGT characters as a uint8 array."""
    rng = np.random.default_rng(seed)
    return _BASES[rng.integers(0, 4, size=n_bases, dtype=np.uint8)]


def write_fasta(path: str, seqs: list[tuple[str, np.ndarray]], width: int = 60) -> None:
    """Write named sequences, `width` bases per line."""
    with open(path, "wb") as f:
        for name, seq in seqs:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def mutate(seq: np.ndarray, err: float, rng: np.random.Generator,
           sub: float = 0.8, ins: float = 0.1) -> np.ndarray:
    """Apply per-base errors: substitution / insertion (before base) / deletion."""
    n = len(seq)
    r = rng.random(n)
    kind = np.zeros(n, dtype=np.int8)          # 0 keep, 1 sub, 2 ins, 3 del
    e = r < err
    r2 = rng.random(n)
    kind[e & (r2 < sub)] = 1
    kind[e & (r2 >= sub) & (r2 < sub + ins)] = 2
    kind[e & (r2 >= sub + ins)] = 3
    out = seq.copy()
    # substitution: rotate to one of the 3 other bases
    s = np.nonzero(kind == 1)[0]
    if len(s):
        idx = np.searchsorted(_BASES, out[s])          # A C G T -> 0..3 (sorted ascii)
        out[s] = _BASES[(idx + rng.integers(1, 4, size=len(s))) % 4]
    reps = np.ones(n, dtype=np.int64)
    reps[kind == 3] = 0
    reps[kind == 2] = 2
    res = np.repeat(out, reps)
    # inserted base = first copy of each doubled position gets a random base
    ipos = np.nonzero(kind == 2)[0]
    if len(ipos):
        starts = np.cumsum(reps) - reps
        res[starts[ipos]] = _BASES[rng.integers(0, 4, size=len(ipos))]
    return res


def simulate_reads(ref: np.ndarray, n_reads: int, read_len: int, err: float, seed: int,
                   n_frac: float = 0.0):
    """Yield (name, bases) reads. Names carry the truth: r<i>_<start>_<strand>."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    starts = rng.integers(0, L - read_len, size=n_reads)
    strands = rng.integers(0, 2, size=n_reads)
    for i in range(n_reads):
        s = int(starts[i])
        frag = ref[s:s + read_len]
        if strands[i]:
            frag = _COMP[frag[::-1]]
        r = mutate(frag, err, rng)
        if n_frac > 0:
            m = rng.random(len(r)) < n_frac
            r = r.copy()
            r[m] = ord("N")
        yield f"r{i}_{s}_{'-' if strands[i] else '+'}", r


def write_reads(path: str, reads, fastq: bool = False) -> int:
    n = 0
    with open(path, "wb") as f:
        for name, seq in reads:
            if fastq:
                f.write(b"@" + name.encode() + b"\n" + seq.tobytes() + b"\n+\n" + b"I" * len(seq) + b"\n")
            else:
                f.write(b">" + name.encode() + b"\n" + seq.tobytes() + b"\n")
            n += 1
    return n


def make_config(outdir: str, ref_bases: int, n_reads: int, read_len: int, err: float,
                ref_seed: int = 12345, read_seed: int = 777, n_seqs: int = 1):
    """Write <outdir>/ref.fa and <outdir>/reads.fa; return their paths."""
    import os
    os.makedirs(outdir, exist_ok=True)
    ref = random_reference(ref_bases, ref_seed)
    bounds = np.linspace(0, ref_bases, n_seqs + 1).astype(np.int64)
    seqs = [(f"chr{k + 1}", ref[bounds[k]:bounds[k + 1]]) for k in range(n_seqs)]
    ref_path = os.path.join(outdir, "ref.fa")
    write_fasta(ref_path, seqs)
    reads_path = os.path.join(outdir, "reads.fa")
    write_reads(reads_path, simulate_reads(ref, n_reads, read_len, err, read_seed))
    return ref_path, reads_path


def edge_reads(ref: np.ndarray, seed: int, n_each: int = 40, read_len: int = 300, err: float = 0.05,
               seq_bounds=None):
    """Reads that exercise the corner cases SURVEY.md section 4 lists: loci hugging offset 0
    with junk prefixes / early insertions (wrapped diagonals, QueryMatch.c:62-67; backward
    clamping, SW.cpp:499-507), loci hugging the reference end (SW.cpp:508-516), N runs
    (Query.c:373-388), sequence-boundary straddlers (AlignOutput.c:129-136), large indels,
    two-piece split reads, very short reads and both strands."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    out = []

    def emit(tag, frag, strand=None):
        if strand is None:
            strand = int(rng.integers(0, 2))
        if strand:
            frag = _COMP[frag[::-1]]
        out.append((f"{tag}_{len(out)}_{'-' if strand else '+'}", np.ascontiguousarray(frag)))

    def junk(n):
        return _BASES[rng.integers(0, 4, size=n)]

    for _ in range(n_each):       # hugging offset 0, junk prefix and/or early insertion
        s = int(rng.integers(0, 40))
        body = mutate(ref[s:s + read_len], err, rng)
        pre = junk(int(rng.integers(0, 40)))
        k = int(rng.integers(16, 60))
        ins = junk(int(rng.integers(0, 12)))
        emit("head", np.concatenate([pre, body[:k], ins, body[k:]]))
    for _ in range(n_each):       # hugging the end
        e = L - int(rng.integers(0, 40))
        body = mutate(ref[e - read_len:e], err, rng)
        emit("tail", np.concatenate([body, junk(int(rng.integers(0, 40)))]))
    for _ in range(n_each):       # N runs
        s = int(rng.integers(0, L - read_len))
        body = mutate(ref[s:s + read_len], err, rng).copy()
        for _k in range(int(rng.integers(1, 4))):
            a = int(rng.integers(0, len(body) - 1))
            body[a:a + int(rng.integers(1, 20))] = ord("N")
        emit("nrun", body)
    if seq_bounds is not None:    # straddle sequence boundaries
        for _ in range(n_each):
            b = int(seq_bounds[int(rng.integers(1, len(seq_bounds) - 1))])
            s = b - int(rng.integers(20, read_len - 20))
            emit("strad", mutate(ref[s:s + read_len], err, rng))
    for _ in range(n_each):       # one large deletion or insertion in the middle
        s = int(rng.integers(0, L - 2 * read_len))
        g = int(rng.integers(5, 95))
        h = read_len // 2
        if rng.integers(0, 2):
            frag = np.concatenate([ref[s:s + h], ref[s + h + g:s + read_len + g]])
        else:
            frag = np.concatenate([ref[s:s + h], junk(g), ref[s + h:s + read_len]])
        emit("indel", mutate(frag, err, rng))
    for _ in range(n_each):       # two-piece split read (possibly inverted second half)
        a = int(rng.integers(0, L - read_len))
        b = int(rng.integers(0, L - read_len))
        h = int(rng.integers(60, read_len - 60))
        second = ref[b:b + read_len - h]
        if rng.integers(0, 2):
            second = _COMP[second[::-1]]
        emit("split", mutate(np.concatenate([ref[a:a + h], second]), err, rng))
    for _ in range(n_each):       # high error
        s = int(rng.integers(0, L - read_len))
        emit("noisy", mutate(ref[s:s + read_len], 0.15, rng))
    for n in (5, 10, 11, 12, 15, 16, 20, 30, 40):   # very short reads (some below K: skipped)
        s = int(rng.integers(0, L - n))
        emit("short", ref[s:s + n].copy())
    emit("junk", junk(read_len))
    return out


def human_like_reference(n_bases: int, n_seqs: int = 24, seed: int = 2024, alu_sites: int | None = None,
                         alu_len: int = 300, alu_div: float = 0.12, n_runs_per_seq: int = 2):
    """BASELINE configs[3]/[4] reference: `n_bases` i.i.d. bases cut into `n_seqs` sequences of chromosome-like unequal
    sizes, with an Alu-like family implanted -- ONE `alu_len`-base consensus copied to `alu_sites` uniformly random
    places (default one per 25 kbp: 124 000 at 3.1 Gbp), every copy on a random strand with its own `alu_div`
    substitutions -- and a few runs of N per sequence (assembly gaps).  The k-mers of the consensus occur tens of
    thousands of times, so the index builder has to down-sample them (Index.c:271-315) and reads that touch a copy
    produce thousands of seed hits per strand.  Returns (bases as one uint8 array, sequence bounds)."""
    rng = np.random.default_rng(seed)
    ref = _BASES[rng.integers(0, 4, size=n_bases, dtype=np.uint8)]
    if alu_sites is None:
        alu_sites = max(8, n_bases // 25_000)
    cons = _BASES[rng.integers(0, 4, size=alu_len, dtype=np.uint8)]
    pos = np.sort(rng.integers(0, n_bases - alu_len, size=alu_sites))
    for a in range(0, alu_sites, 1 << 16):                           # in slabs: bounded temporaries
        p = pos[a:a + (1 << 16)]
        m = len(p)
        copies = np.tile(cons, (m, 1))
        sub = rng.random((m, alu_len)) < alu_div
        idx = np.searchsorted(_BASES, copies[sub])
        copies[sub] = _BASES[(idx + rng.integers(1, 4, size=int(sub.sum()))) % 4]
        flip = rng.integers(0, 2, size=m).astype(bool)
        copies[flip] = _COMP[copies[flip][:, ::-1]]
        ref[(p[:, None] + np.arange(alu_len)[None, :]).ravel()] = copies.ravel()
    w = np.linspace(2.0, 0.5, n_seqs)
    bounds = np.concatenate([[0], np.cumsum(w / w.sum() * n_bases)]).astype(np.int64)
    bounds[-1] = n_bases
    for k in range(n_seqs):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        for _ in range(n_runs_per_seq):
            ln = int(rng.integers(50, 20_000))
            if hi - lo > 4 * ln:
                a = int(rng.integers(lo, hi - ln))
                ref[a:a + ln] = ord("N")
    return ref, bounds


def sv_reads(ref: np.ndarray, n_reads: int, read_len: int, seed: int, err_lo: float = 0.02, err_hi: float = 0.05):
    """BASELINE configs[3] reads: `read_len`-base reads (10 kbp) of a donor that differs from the reference by one to
    three structural variants inside the read -- a deletion (50-2000 bases missing), an inversion (300-3000 bases
    reverse-complemented in place), a novel 300-base insertion, or a translocated piece from elsewhere -- plus 2-5 %
    per-base errors: the split-read sets -OQC / -FBS are made for."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    for i in range(n_reads):
        s = int(rng.integers(0, L - 2 * read_len))
        w = ref[s:s + read_len + 6000].copy()
        kinds = []
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, 4))
            at = int(rng.integers(500, read_len - 500))
            if k == 0:
                n = int(rng.integers(50, 2000)); w = np.concatenate([w[:at], w[at + n:]]); kinds.append(f"del{n}")
            elif k == 1:
                n = int(rng.integers(300, 3000)); w = np.concatenate([w[:at], _COMP[w[at:at + n][::-1]], w[at + n:]]); kinds.append(f"inv{n}")
            elif k == 2:
                w = np.concatenate([w[:at], _BASES[rng.integers(0, 4, size=300)], w[at:]]); kinds.append("ins300")
            else:
                n = int(rng.integers(500, 3000)); b = int(rng.integers(0, L - n))
                w = np.concatenate([w[:at], ref[b:b + n], w[at:]]); kinds.append(f"tra{n}")
        w = w[:read_len]
        strand = int(rng.integers(0, 2))
        if strand:
            w = _COMP[w[::-1]]
        w = w[w != 0] if (w == 0).any() else w                      # (_COMP of a non-ACGTN byte: none here)
        yield f"sv{i}_{s}_{'-' if strand else '+'}_{'_'.join(kinds)}", mutate(w, float(rng.uniform(err_lo, err_hi)), rng)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("outdir")
    ap.add_argument("--ref-bases", type=int, default=10_000_000)
    ap.add_argument("--reads", type=int, default=10_000)
    ap.add_argument("--len", type=int, default=1000)
    ap.add_argument("--err", type=float, default=0.02)
    ap.add_argument("--seqs", type=int, default=1)
    ap.add_argument("--ref-seed", type=int, default=12345)
    ap.add_argument("--read-seed", type=int, default=777)
    a = ap.parse_args()
    print(make_config(a.outdir, a.ref_bases, a.reads, a.len, a.err, a.ref_seed, a.read_seed, a.seqs))


def n_rich_reference(seed: int = 3):
    """Five sequences (50 001, 12, 30 007, 999 and 20 000 bases) with runs of N of 1-40 bases about every 300 bases, one
    sequence starting and one ending with N, one shorter than most word lengths: what the index builder's walk over
    non-ACGT codes (Index.c:95-128) has to get right for every -L / -S."""
    rng = np.random.default_rng(seed)
    seqs = []
    for k, L in enumerate([50_001, 12, 30_007, 999, 20_000]):
        s = random_reference(L, 50 + k).copy()
        for _ in range(L // 300):
            p = int(rng.integers(0, L))
            n = int(rng.choice([1, 1, 2, 3, 5, 8, 13, 40]))
            s[p:p + n] = ord("N")
        if k == 2:
            s[:7] = ord("N")
        if k == 4:
            s[-5:] = ord("N")
        seqs.append((f"s{k}", s))
    return seqs
