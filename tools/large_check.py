#!/usr/bin/env python
"""Large-reference end-to-end check (not part of pytest: needs ~40 GB of RAM/disk and minutes).

Builds a multi-sequence synthetic reference of --gbp giga-bases (offsets beyond 2^31 exercise the
unsigned 32-bit paths), indexes it ON THE DEVICE, writes the index in the reference's format, aligns
simulated reads with the product host program and with the unmodified reference binary
(oracle/_ref/yaha) and compares the SAM byte for byte (all lines except @PG).

    python tools/large_check.py --gbp 2.5 --reads 4000 --len 1000 --err 0.05 -- -H 650 -MD 50
"""
import argparse
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yaha_b200                      # noqa: E402
from yaha_b200 import refio, synth    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbp", type=float, default=1.0)
    ap.add_argument("--seqs", type=int, default=12)
    ap.add_argument("--reads", type=int, default=4000)
    ap.add_argument("--len", type=int, default=1000)
    ap.add_argument("--err", type=float, default=0.05)
    ap.add_argument("--sv", type=int, default=0, help="additionally simulate this many chimeric split reads (2-5 pieces, some inverted)")
    ap.add_argument("--dir", default="/tmp/yaha_b200_large")
    ap.add_argument("flags", nargs="*")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    n = int(a.gbp * 1e9)
    t = time.time()
    ref = synth.random_reference(n, 2024)
    bounds = np.linspace(0, n, a.seqs + 1).astype(np.int64)
    seqs = [(f"chr{k + 1}", ref[bounds[k]:bounds[k + 1]]) for k in range(a.seqs)]
    nib_path = os.path.join(a.dir, "ref.nib2")
    with open(nib_path, "wb") as f:
        f.write(refio.build_nib2(seqs))
    nib = refio.load_nib2(nib_path)
    print(f"reference {n} bases in {a.seqs} sequences, maxROff {nib.max_roff}, {time.time() - t:.1f}s", flush=True)
    t = time.time()
    al = yaha_b200.Aligner(nib, None, yaha_b200.Params.defaults(word_len=15), device=0)
    idx = al.download_index()
    al.close()
    idx_path = os.path.join(a.dir, refio.index_file_name("ref", 15, 1, 65525))
    refio.write_index(idx_path, idx)
    print(f"device index build + write: {len(idx.roa)} entries, {time.time() - t:.1f}s", flush=True)
    reads = list(synth.simulate_reads(ref, a.reads, a.len, a.err, 99))
    # a few reads hugging the very end of the reference (offsets close to maxROff) and its start
    rng = np.random.default_rng(3)
    for k in range(50):
        s = n - a.len - int(rng.integers(0, 30))
        reads.append((f"tail{k}", synth.mutate(ref[s:s + a.len], a.err, rng)))
        s = int(rng.integers(0, 30))
        reads.append((f"head{k}", np.concatenate([synth._BASES[rng.integers(0, 4, size=20)], synth.mutate(ref[s:s + a.len], a.err, rng)])))
    for k in range(a.sv):                        # BASELINE configs[3] shape: long reads spanning rearrangements
        pieces = []
        for _ in range(int(rng.integers(2, 6))):
            s = int(rng.integers(0, n - 4000)); ln = int(rng.integers(300, 3500))
            p = ref[s:s + ln]
            if rng.integers(0, 3) == 0:
                p = synth._COMP[p[::-1]]
            pieces.append(p)
        reads.append((f"sv{k}", synth.mutate(np.concatenate(pieces), 0.03, rng)))
    q = os.path.join(a.dir, "reads.fa")
    synth.write_reads(q, reads)
    del ref
    host = os.path.join(ROOT, "yaha_b200", "yaha_b200_host")
    refbin = os.path.join(ROOT, "oracle", "_ref", "yaha")
    t = time.time()
    subprocess.run([host, "-x", idx_path, "-q", q, "-osh", os.path.join(a.dir, "mine.sam"), "-t", str(os.cpu_count())] + a.flags, check=True)
    t_mine = time.time() - t
    t = time.time()
    subprocess.run([refbin, "-x", idx_path, "-q", q, "-osh", os.path.join(a.dir, "ref.sam"), "-t", "1"] + a.flags, check=True)
    t_ref = time.time() - t
    mine = [l for l in open(os.path.join(a.dir, "mine.sam")) if not l.startswith("@PG")]
    want = [l for l in open(os.path.join(a.dir, "ref.sam")) if not l.startswith("@PG")]
    same = mine == want
    print(f"records {len(want)}  identical {same}  host wall {t_mine:.1f}s (incl. index upload)  reference -t 1 wall {t_ref:.1f}s", flush=True)
    if not same:
        for i, (x, y) in enumerate(zip(mine, want)):
            if x != y:
                print("first difference at line", i, "\n", x[:300], "\n", y[:300])
                break
        sys.exit(1)


if __name__ == "__main__":
    main()
