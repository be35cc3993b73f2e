#!/usr/bin/env python
"""Differential fuzzing of the product's host program + shared per-read sources against the UNMODIFIED reference, on the CPU.

The per-read steps the kernels run (csrc/form_clumps.h, prepare_clumps.h, assemble_clumps.h, finish_reads.h) are the very sources
the mock of the ABI (tests/mock/mock_abi.c) compiles for the host, so a disagreement found here is a disagreement of the device
path.  Every case draws a reference (random sequence + repeat families, tandem repeats, homopolymers, N runs, several sequences),
an index geometry (-L/-S/-H), a read mix (plain, edge cases, chimeras, multi-piece, pieces of repeats) and a flag set, runs
oracle/_ref/yaha -t 1 and tests/_build/yaha_host_mock, and compares the SAM line by line (all lines but @PG).

Build-container tool (needs oracle/_ref, i.e. /root/reference at build time); failures are kept under --keep for turning into
goldens (tests/golden/make_golden.py).   usage: tools/fuzz_parity.py --seeds 0:200 [--jobs 6] [--keep /tmp/fuzz_fail] [--heavy]
On a GPU box the same cases run against the CUDA library: --binary yaha_b200/yaha_b200_host (oracle/_ref/yaha travels with the repo).
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from yaha_b200 import synth  # noqa: E402
import hostcases as H  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "yaha")
MOCK = os.path.join(ROOT, "tests", "_build", "yaha_host_mock")


def draw_reference(rng, heavy=False, manyseq=False):
    n = int(rng.integers(40_000, 160_000))
    ref = synth.random_reference(n, int(rng.integers(1, 1 << 30))).copy()
    # repeat families: an element pasted many times with divergence
    for _ in range(int(rng.integers(2, 5)) if heavy else int(rng.integers(0, 4))):
        ln = int(rng.integers(60, 700))
        elem = synth._BASES[rng.integers(0, 4, size=ln)]
        div = float(rng.choice([0.0, 0.01, 0.05, 0.12]))
        for _c in range(int(rng.integers(40, 300)) if heavy else int(rng.integers(3, 60))):
            p = int(rng.integers(0, n - ln - 1))
            cp = synth.mutate(elem, div, rng)[:ln]
            if rng.integers(0, 2):
                cp = synth._COMP[cp[::-1]]
            ref[p:p + len(cp)] = cp
    # tandem repeats and homopolymers
    for _ in range(int(rng.integers(0, 6))):
        unit = synth._BASES[rng.integers(0, 4, size=int(rng.integers(1, 40)))]
        reps = int(rng.integers(3, 60))
        t = np.tile(unit, reps)[:2000]
        p = int(rng.integers(0, n - len(t) - 1))
        ref[p:p + len(t)] = t
    # segmental duplication (exact, long)
    for _ in range(int(rng.integers(0, 3))):
        ln = int(rng.integers(500, 4000))
        a, b = int(rng.integers(0, n - ln)), int(rng.integers(0, n - ln))
        ref[b:b + ln] = ref[a:a + ln]
    # N runs
    for _ in range(int(rng.integers(0, 4))):
        p = int(rng.integers(0, n - 300))
        ref[p:p + int(rng.integers(1, 200))] = ord("N")
    k = int(rng.integers(1, 5))
    if manyseq:
        k = int(rng.integers(40, 400))                                # hundreds of short sequences: boundaries everywhere
    bounds = sorted(set([0, n] + [int(x) for x in rng.integers(100 if manyseq else 1000, n - (100 if manyseq else 1000), size=k - 1)]))
    return ref, bounds


def draw_reads(rng, ref, bounds, heavy=False, chars=False):
    import make_golden as MG
    L = len(ref)
    reads = []
    kind = int(rng.integers(0, 6))
    nplain = int(rng.integers(10, 40)) if heavy else int(rng.integers(20, 120))
    for i in range(nplain):
        ln = int(rng.choice([1000, 2500, 6000, 12000, 30000])) if heavy else int(rng.choice([25, 40, 80, 150, 300, 600, 1500, 4000]))
        ln = min(ln, L - 10)
        s = int(rng.integers(0, L - ln))
        err = float(rng.choice([0.0, 0.01, 0.03, 0.08, 0.15]))
        r = synth.mutate(ref[s:s + ln], err, rng)
        if rng.integers(0, 2):
            r = synth._COMP[r[::-1]]
        reads.append((f"p{i}_{s}", r))
    if kind in (0, 1, 5):
        reads += synth.edge_reads(ref, int(rng.integers(1, 1 << 30)), n_each=int(rng.integers(2, 8)), read_len=int(rng.choice([150, 300, 500])),
                                  err=float(rng.choice([0.02, 0.05, 0.1])), seq_bounds=bounds if len(bounds) > 2 else None)
    if kind in (1, 2, 5) and L > 8000:
        reads += MG.chimera_reads(ref, int(rng.integers(1, 1 << 30)), n_reads=int(rng.integers(5, 40)))
    if kind in (2, 3, 4, 5):
        reads += MG.multi_reads(ref, int(rng.integers(1, 1 << 30)), n=int(rng.integers(10, 80)))
    # character-level noise: lower case, IUPAC codes, stray symbols (Math.c:141-157: the code table of the reader)
    if chars:
        noisy = []
        for name, r in reads:
            r = np.array(r, dtype=np.uint8, copy=True)
            u = rng.random()
            if u < 0.10 and len(r):
                a = int(rng.integers(0, len(r))); b = int(rng.integers(a, len(r) + 1))
                seg = r[a:b]; up = (seg >= 65) & (seg <= 90); seg[up] += 32
            elif u < 0.16 and len(r):
                k = int(rng.integers(1, 6))
                pos = rng.integers(0, len(r), size=k)
                r[pos] = np.frombuffer(b"RYKMSWBDHVNXryn.-*", dtype=np.uint8)[rng.integers(0, 18, size=k)]
            noisy.append((name, r))
        reads = noisy
    # order shuffled so that batches mix kinds
    order = rng.permutation(len(reads))
    return [reads[int(k)] for k in order]


def write_queries(path, reads, fastq, rng, weird):
    """The query file.  weird: the reader's corner cases (Query.c:102-228) -- descriptions and record markers inside id lines,
    multi-line sequences, blank lines, CRLF, empty records, missing final newline, random qualities, '+' lines with names."""
    if not weird:
        return synth.write_reads(path, reads, fastq=fastq)
    eol = b"\r\n" if rng.integers(0, 6) == 0 else b"\n"
    out = bytearray()
    last_len = 1
    for name, seq in reads:
        seq = np.asarray(seq, dtype=np.uint8)
        nm = name.encode()
        u = int(rng.integers(0, 10))
        if u == 0: nm += b" some description"
        elif u == 1: nm += b"\tx=1"
        elif u == 2: nm += b" >inner @at +plus"
        elif u == 3: nm += b">" 
        if rng.integers(0, 60) == 0:
            seq = seq[:0]                                             # empty record
        if rng.integers(0, 150) == 0 and len(seq):
            seq = np.tile(seq, 32100 // len(seq) + 1)[:int(rng.integers(32001, 32100))]   # longer than the reader's 32 000 bases
        if rng.integers(0, 100) == 0:
            nm = nm + b"_" + b"x" * int(rng.integers(190, 230))        # id around the 200-character cut
        last_len = len(seq)
        if fastq:
            q = rng.integers(35, 127, size=len(seq)).astype(np.uint8)  # (no '!' / '"' / '#': nothing special, just printable)
            if len(q) and rng.integers(0, 4) == 0: q[0] = ord("@")
            if len(q) > 1 and rng.integers(0, 4) == 0: q[int(rng.integers(0, len(q)))] = ord("+")
            plus = b"+" + (nm if rng.integers(0, 4) == 0 else b"")
            if rng.integers(0, 40) == 0 and len(q) > 2: q = q[:-1]      # quality shorter than the sequence: record skipped
            sb, qb = seq.tobytes(), q.tobytes()
            if rng.integers(0, 25) == 0 and len(sb) > 20:                # multi-line FASTQ
                h = len(sb) // 2
                sb = sb[:h] + eol + sb[h:]
                qb = qb[:h] + eol + qb[h:]
            out += b"@" + nm + eol + sb + eol + plus + eol + qb + eol
        else:
            out += b">" + nm + eol
            w = int(rng.choice([0, 0, 60, 70, 13])) or max(1, len(seq))
            if rng.integers(0, 8) == 0: w = int(rng.integers(1, 200))
            for a in range(0, len(seq), w):
                line = seq[a:a + w].tobytes()
                if rng.integers(0, 400) == 0 and len(line) > 4: line = line[:2] + b">" + line[2:]   # a marker inside a sequence line
                out += line + eol
                if rng.integers(0, 300) == 0: out += eol               # blank line inside a record
            if rng.integers(0, 30) == 0: out += eol
    if rng.integers(0, 5) == 0 and out.endswith(eol) and last_len > 0:
        del out[-len(eol):]
    open(path, "wb").write(bytes(out))
    return len(reads)


FLAG_POOL = H.FLAG_SWEEP + [[], [], [], ["-OQC", "N"], ["-FBS", "Y"], ["-FBS", "Y", "-PRL", "0.3", "-PSS", "0.3"], ["-BW", "10", "-G", "100"],
                            ["-MGDP", "3", "-BP", "3"], ["-M", "18", "-P", "0.6"], ["-X", "15", "-BW", "15", "-G", "150"]]


def one_case(args):
    seed, keep, heavy, binary, wordlens, chars, weird, manyseq, from_stdin = args
    rng = np.random.default_rng(seed)
    tmp = tempfile.mkdtemp(prefix=f"fuzz{seed}_")
    try:
        ref, bounds = draw_reference(rng, heavy, manyseq)
        synth.write_fasta(tmp + "/ref.fa", [(f"chr{k + 1} desc", ref[bounds[k]:bounds[k + 1]]) for k in range(len(bounds) - 1)],
                          width=int(rng.choice([50, 60, 70])))
        Lw = int(rng.choice(wordlens))
        Sk = int(rng.choice([1, 1, 1, 1, 2, 3]))
        Hh = int(rng.choice([65525, 65525, 650, 40, 8]))
        gen = [REF, "-g", "ref.fa", "-L", str(Lw), "-S", str(Sk)] + (["-H", str(Hh)] if Hh != 65525 else [])
        subprocess.check_call(gen, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        idx = [f for f in os.listdir(tmp) if f.startswith("ref.X")]
        assert len(idx) == 1, idx
        fastq = bool(rng.integers(0, 5) == 0)
        reads = draw_reads(rng, ref, bounds, heavy, chars)
        qf = "reads.fq" if fastq else "reads.fa"
        write_queries(tmp + "/" + qf, reads, fastq, rng, weird)
        flags = list(FLAG_POOL[int(rng.integers(0, len(FLAG_POOL)))])
        if rng.integers(0, 3) == 0:
            more = FLAG_POOL[int(rng.integers(0, len(FLAG_POOL)))]
            have = set(flags[0::2])
            for a, b in zip(more[0::2], more[1::2]):
                if a not in have:
                    flags += [a, b]
        outflag = str(rng.choice(["-osh", "-osh", "-osh", "-oss", "-o8"]))
        base = ["-x", idx[0]] + ([] if from_stdin else ["-q", qf])            # (no -q: queries from standard input, Main.c:173-178)
        r = subprocess.run([REF] + base + [outflag, "want.out", "-t", "1"] + flags, cwd=tmp, capture_output=True, text=True, timeout=900,
                           stdin=open(tmp + "/" + qf, "rb") if from_stdin else subprocess.DEVNULL)
        if r.returncode != 0:
            return seed, f"skip (reference failed rc={r.returncode}: {r.stderr.strip()[-160:]!r}): " + " ".join(flags), None
        host = ["-t", str(int(rng.integers(1, 4))), "-batch", str(int(rng.choice([7, 33, 100, 5000]))), "-pipes", str(int(rng.integers(1, 3)))]
        if rng.integers(0, 5) == 0 and binary == MOCK:             # (the mock pretends to have any number of devices)
            host += ["-gpus", "2"]
        env = dict(os.environ)
        if rng.integers(0, 4) == 0:
            env["YA_FUSED"] = "0"
        m = subprocess.run([binary] + base + [outflag, "got.out"] + host + flags, cwd=tmp, capture_output=True, text=True, timeout=1800, env=env,
                           stdin=open(tmp + "/" + qf, "rb") if from_stdin else subprocess.DEVNULL)
        desc = f"L{Lw} S{Sk} H{Hh} {qf} {outflag} {' '.join(flags)} | host {' '.join(host)} fused={env.get('YA_FUSED', '1')} reads={len(reads)} ref={len(ref)}"
        if m.returncode != 0:
            bad = "mock failed rc=%d: %s" % (m.returncode, m.stderr[-600:])
        else:
            got, want = H.sam_lines(open(tmp + "/got.out").read()), H.sam_lines(open(tmp + "/want.out").read())
            bad = None
            if got != want:
                k = next((i for i, (x, y) in enumerate(zip(got, want)) if x != y), min(len(got), len(want)))
                bad = f"line {k} of {len(got)}/{len(want)}:\n  got  {got[k][:300] if k < len(got) else None}\n  want {want[k][:300] if k < len(want) else None}"
        if bad and keep:
            dst = os.path.join(keep, f"seed{seed}")
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(tmp, dst, ignore=shutil.ignore_patterns("ref.X*", "ref.nib2"))     # (the index is rebuilt by `yaha -g`: CASE.txt has -L/-S/-H)
            open(os.path.join(dst, "CASE.txt"), "w").write(desc + "\n" + bad + "\n")
        return seed, desc, bad
    except subprocess.TimeoutExpired as e:
        return seed, "timeout " + str(e.cmd[:1]), None
    except Exception as e:                                  # generator corner (e.g. a reference too short for a read kind)
        return seed, "skip (generator): " + repr(e)[:200], None
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:50")
    ap.add_argument("--jobs", type=int, default=6)
    ap.add_argument("--keep", default="/tmp/fuzz_fail")
    ap.add_argument("--heavy", action="store_true", help="repeat families with 40-300 copies, reads of 1-30 kbp")
    ap.add_argument("--binary", default=MOCK, help="host program under test (default: the CPU mock build; on a GPU box: yaha_b200/yaha_b200_host)")
    ap.add_argument("--chars", action="store_true", help="lower case, IUPAC codes and stray symbols in a quarter of the reads")
    ap.add_argument("--weird", action="store_true", help="query files with the reader's corner cases (multi-line, CRLF, markers in id lines, random qualities ...)")
    ap.add_argument("--manyseq", action="store_true", help="references of 40-400 short sequences")
    ap.add_argument("--stdin", action="store_true", help="queries through standard input (no -q)")
    ap.add_argument("--wordlens", default="11,11,12,13,15", help="-L values drawn from (an -L 15 index is a 4.3 GB file)")
    a = ap.parse_args()
    a.binary = os.path.abspath(a.binary)
    assert os.access(a.binary, os.X_OK) or a.binary == MOCK, a.binary
    lo, hi = (int(x) for x in a.seeds.split(":"))
    wl = [int(x) for x in a.wordlens.split(",")]
    os.makedirs(a.keep, exist_ok=True)
    if a.binary == MOCK:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "mock"), "SAN="])
    nbad = nskip = 0
    with ProcessPoolExecutor(a.jobs) as ex:
        for seed, desc, bad in ex.map(one_case, [(s, a.keep, a.heavy, a.binary, wl, a.chars, a.weird, a.manyseq, a.stdin) for s in range(lo, hi)]):
            print(("FAIL" if bad else "ok  "), seed, desc, flush=True)
            nskip += desc.startswith(("skip", "timeout"))
            if bad:
                nbad += 1
                print(bad, flush=True)
    print(f"{hi - lo} cases, {hi - lo - nskip} compared, {nbad} failed")
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
