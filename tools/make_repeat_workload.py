#!/usr/bin/env python
"""Repeat-rich local workload (BASELINE configs[3] flavour at small scale): an 8 Mbp reference with 3 000 diverged copies of
three 300 bp Alu-like elements and 40 tandem arrays, 2 500 x 1000 bp reads at 5 %.  Writes ref.fa / ref.nib2 / the index /
reads.fa into the current directory; compare `oracle/_ref/yaha -t 1` with the host program (or tests/_build/yaha_host_mock,
no GPU needed) on it.  r01: SAM identical in all three modes (device clumps + phase 1, YA_HOST_PREP=1, YA_HOST_CLUMPS=1);
632 DP jobs per read; 22 DP rounds for one 2 500-read batch since clumps that split are scored in child fibers (708 before)."""
import sys, os, time
sys.path.insert(0, "/root/repo")
from yaha_b200 import refio, synth
import numpy as np
rng = np.random.default_rng(5)
nb = 8_000_000
ref = synth.random_reference(nb, 4242).copy()
# Alu-like repeats: 3000 diverged copies of three 300 bp elements, plus tandem arrays
elems = [synth.random_reference(300, 100 + k) for k in range(3)]
for _ in range(3000):
    e = elems[int(rng.integers(0, 3))]
    cp = synth.mutate(e, 0.08, rng)[:300]
    pos = int(rng.integers(1000, nb - 2000))
    ref[pos:pos + len(cp)] = cp
for _ in range(40):                      # tandem arrays of a 31 bp unit (many fragments per region)
    unit = synth.random_reference(31, int(rng.integers(0, 1 << 30)))
    pos = int(rng.integers(1000, nb - 5000))
    arr = np.tile(unit, 60)
    m = synth.mutate(arr, 0.03, rng)[:len(arr)]
    ref[pos:pos + len(m)] = m
synth.write_fasta("ref.fa", [("chrR", ref)])
open("ref.nib2", "wb").write(refio.build_nib2([("chrR", ref)]))
nib = refio.load_nib2("ref.nib2")
idx = refio.build_index(nib, 15)
p = refio.index_file_name("ref", 15, 1, 65525)
open(p, "wb").write(idx)
reads = list(synth.simulate_reads(ref, 2500, 1000, 0.05, 99))
synth.write_reads("reads.fa", reads)
print("done", p)
