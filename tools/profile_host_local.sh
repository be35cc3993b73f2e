#!/bin/bash
# Worker time per read of the host program WITHOUT a GPU: the product's host sources on the oracle-backed mock of the ABI
# (tests/mock), built -O2 like the product, on a local cfg3-shaped (500 bp, 10 %) and cfg2-shaped (100 bp, 5 %) workload over a
# 20 Mbp reference.  Prints the phase cycle counters (YAHA_B200_PROF=1) and the worker wall per read, minimum of N runs.
# DESIGN.md section 8 quotes these numbers.  usage: tools/profile_host_local.sh [N] [workdir]
set -e
N=${1:-5}; W=${2:-/tmp/yaha_b200_local}; ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $W/bin
make -s -C $ROOT/tests/mock OPT="-O2" SAN= OUT=$W/bin
cd $W
if [ ! -f ref.X15_01_65525S ]; then
PYTHONPATH=$ROOT python - <<'PY'
from yaha_b200 import refio, synth
ref = synth.random_reference(20_000_000, 12345)
open("ref.nib2", "wb").write(refio.build_nib2([("chr1", ref)]))
open(refio.index_file_name("ref", 15, 1, 65525), "wb").write(refio.build_index(refio.load_nib2("ref.nib2"), 15))
synth.write_reads("reads500.fa", list(synth.simulate_reads(ref, 5000, 500, 0.10, 777)))
synth.write_reads("reads100.fa", list(synth.simulate_reads(ref, 8000, 100, 0.05, 6)))
PY
fi
for spec in "reads500.fa -BW 10 -G 100" "reads100.fa"; do
  for i in $(seq $N); do
    YAHA_B200_PROF=1 taskset -c 1 bin/yaha_host_mock -x ref.X15_01_65525S -q $spec -osh out.sam -t 1 -passes 2 2>&1 | grep -E '"pass"|prof' | tail -3
  done | python3 -c "
import sys, json, re
wall, rows = [], []
for l in sys.stdin:
    if l.startswith('{'): d = json.loads(l); wall.append(d['host_wall_s'] / d['reads'] * 1e6)
    else: rows.append(l.strip())
print('$spec'.split()[0], 'worker wall per read, us: min %.2f median %.2f' % (min(wall), sorted(wall)[len(wall) // 2]))
print('  (counters of the last run, Mcycles over 2 passes)', ' | '.join(rows[-2:]))"
done
