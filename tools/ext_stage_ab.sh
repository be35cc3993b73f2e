#!/bin/bash
# A/B of the extension kernel with and without reference-window staging by cp.async.bulk (YA_EXT_STAGE), on the GPU box after
# bench.py filled its cache: CUDA-event time of the bulk extension launches for the whole-shard batch (run C sizes) and for
# 5000-read batches, three repetitions each; and one `ncu --set full` capture per variant.
D=${YAHA_BENCH_CACHE:-/tmp}/yaha_b200_bench_iid100; X=$(ls $D/ref.X15_01_* | head -1); Q=$D/reads_rank0.fa
H=yaha_b200/yaha_b200_host; O=gpurun_out; mkdir -p $O
for B in 20000 5000; do
  for S in 0 1 0 1 0 1; do
    YA_EXT_STAGE=$S timeout 120 $H -x $X -q $Q -osh /tmp/ab.sam -t 4 -batch $B -pipes 1 -passes 8 -replay -BW 10 -G 100 2>&1 | grep '"pass"' | tail -5 | python3 -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
c=sum(x['ext_cells'] for x in r); ms=sum(x['dev_ms_ext'] for x in r); n=sum(x['ext_launches'] for x in r)
print('batch %5d stage %d: %d bulk ext launches, %.1f us per launch, %.1f GCUPS' % ($B, $S, n, 1e3*ms/max(n,1), c/max(ms,1e-9)/1e6))"
  done
done
for S in 0 1; do
  YA_EXT_STAGE=$S timeout -s KILL 300 ncu --set full --clock-control none -k regex:dp_ext_packed -c 1 -f -o $O/prof_ext_stage$S \
      $H -x $X -q $Q -osh /tmp/ab.sam -t 4 -batch 20000 -pipes 1 -BW 10 -G 100 > $O/ncu_ext_stage$S.log 2>&1
  ncu -i $O/prof_ext_stage$S.ncu-rep --page raw --csv > $O/prof_ext_stage$S.raw.csv 2>/dev/null; rm -f $O/prof_ext_stage$S.ncu-rep
done
