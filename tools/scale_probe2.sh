#!/bin/bash
# Follow-up of scale_probe.sh: (a) do the processes really run their passes at the same time (wall clock of pass 8)?  (b) does
# the fast process follow barrier rank 0 or device 0 (rank r on device (r+3)%N)?  (c) every process sees only its own GPU.
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
D=/tmp/yaha_b200_bench_iid100; X=$(ls $D/ref.X15_01_* | head -1); H=yaha_b200/yaha_b200_host
N=${1:-8}; CORES=$(nproc); PER=$((CORES / N))
for r in $(seq 1 $((N - 1))); do cp $D/reads_rank0.fa $D/reads_rank$r.fa 2>/dev/null; done
summ() { python3 -c "
import sys, json
L=sys.stdin.read().splitlines()
r=[json.loads(l) for l in L if l.startswith('{\"pass\"')][6:]
t8=[l for l in L if l.startswith('pass 8 started')]
med=lambda k: sorted(x[k] for x in r)[len(r)//2]*1e3
print('$1: step %.2f ms  parse %.2f  write %.2f  upload %.2f | %s' % (med('align_s'), med('read_parse_s'), med('write_s'), med('upload_s'), t8[0][-28:] if t8 else ''))"; }
for mode in shifted visible; do
  echo "== $N processes together, $mode"
  B=$D/probe2_barrier_$mode; rm -rf $B; mkdir -p $B
  for r in $(seq 0 $((N - 1))); do
    if [ $mode = shifted ]; then DEV=$(( (r + 3) % N )); VIS=""; else DEV=0; VIS="CUDA_VISIBLE_DEVICES=$r"; fi
    ( env $VIS YA_PASS_CLOCK=1 YA_START_BARRIER=$B:$r:$N YA_NAP_US=50 $H -x $X -q $D/reads_rank$r.fa -osh $D/probe_$r.sam -t $PER -dev $DEV -batch 5000 -pipes 4 -passes 16 -BW 10 -G 100 2>&1 | summ "rank$r dev$DEV($mode)" ) &
  done
  wait
done
