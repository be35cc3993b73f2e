#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
for tb in thread warp; do
for cfg in "20000 1 20000" "2500 8 0"; do
  set -- $cfg
  YA_TB=$tb YA_COALESCE_US=$3 yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/tb_$tb.sam -t 16 -BW 10 -G 100 -batch $1 -pipes $2 -passes 14 -replay 2>&1 | grep '"pass"' | tail -6 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v)
print('tb=$tb batch=$1 pipes=$2', int(sum(x['reads_per_s'] for x in v)/n), 'dev_ms_tb', round(sum(x['dev_ms_traceback'] for x in v)/n,3), 'dev_ms_dp', round(sum(x['dev_ms_dp'] for x in v)/n,3), 'dp_wall_ms', round(1e3*sum(x['dp_wall_s'] for x in v)/n,2), 'rounds', v[-1]['dp_rounds'])"
done; done
cmp /tmp/tb_thread.sam /tmp/tb_warp.sam && echo SAM_IDENTICAL
