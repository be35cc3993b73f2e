#!/bin/bash
# Sweep of batch size / pipelines / worker threads of the host program on the cfg3 bench workload (run after bench.py has
# filled its cache directory on the GPU box).  Prints the median align_s of the last passes for every configuration.
# usage: tools/sweep_fused.sh [cache_dir] [reads file] [extra host flags...]
D=${1:-/tmp/yaha_b200_bench_iid100}; Q=${2:-$D/reads_rank0.fa}; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
IDX=$(ls $D/ref.X15_01_* | head -1)
for spec in "20000 1" "10000 2" "5000 2" "5000 4" "2500 4" "2500 8" "1250 8" "5000 8"; do
  set -- $spec; B=$1; P=$2
  for T in 2 4 16; do
    for mode in "" "-replay"; do
      $ROOT/yaha_b200/yaha_b200_host -x $IDX -q $Q -osh $D/sweep.sam -t $T -batch $B -pipes $P -passes 14 $mode -BW 10 -G 100 2>&1 >/dev/null | grep '"pass"' | tail -8 | python3 -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
a=sorted(x['align_s'] for x in r)
print('batch %5d pipes %d threads %2d %-8s median %.2f ms  min %.2f ms  -> %.2f M reads/s | dev ms: seed %.2f dp %.2f tb %.2f finish %.2f | handed back %d launches %d' % ($B, $P, $T, '$mode' or 'e2e', a[len(a)//2]*1e3, a[0]*1e3, r[0]['reads']/a[len(a)//2]/1e6, r[-1]['dev_ms_seed'], r[-1]['dev_ms_dp'], r[-1]['dev_ms_traceback'], r[-1]['dev_ms_finish'], r[-1]['reads_handed_back'], r[-1]['launches']))"
    done
  done
done
