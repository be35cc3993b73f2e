#!/bin/bash
# On an 8-GPU box: where do the ranks of a multi-process run lose time against a single process?  (i) one process alone on
# device 5; (ii) eight processes, one per device, started together (YA_START_BARRIER); (iii) the same, each pinned to its own
# cores.  Prints per process the median step, parse and write time of the e2e job.
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
D=/tmp/yaha_b200_bench_iid100; X=$(ls $D/ref.X15_01_* | head -1); H=yaha_b200/yaha_b200_host
N=${1:-8}; CORES=$(nproc); PER=$((CORES / N))
for r in $(seq 1 $((N - 1))); do cp $D/reads_rank0.fa $D/reads_rank$r.fa 2>/dev/null; done
summ() { python3 -c "
import sys, json
r=[json.loads(l) for l in sys.stdin if l.startswith('{\"pass\"')][6:]
med=lambda k: sorted(x[k] for x in r)[len(r)//2]*1e3
print('$1: step %.2f ms  parse %.2f  write %.2f  upload %.2f  dp(4 pipes) %.2f' % (med('align_s'), med('read_parse_s'), med('write_s'), med('upload_s'), med('dp_wall_s')))"; }
echo "== one process alone on device 5"
$H -x $X -q $D/reads_rank0.fa -osh $D/probe_5.sam -t $PER -dev 5 -batch 5000 -pipes 4 -passes 16 -BW 10 -G 100 2>&1 | summ dev5-alone
for mode in free pinned; do
  echo "== $N processes together, $mode"
  B=$D/probe_barrier_$mode; rm -rf $B; mkdir -p $B
  for r in $(seq 0 $((N - 1))); do
    PIN=""; [ $mode = pinned ] && PIN="taskset -c $((r * PER))-$((r * PER + PER - 1))"
    ( YA_START_BARRIER=$B:$r:$N YA_NAP_US=50 $PIN $H -x $X -q $D/reads_rank$r.fa -osh $D/probe_$r.sam -t $PER -dev $r -batch 5000 -pipes 4 -passes 16 -BW 10 -G 100 2>&1 | summ "dev$r" ) &
  done
  wait
done
