#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { # label, extra host args...
  lab=$1; shift
  yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 "$@" > /tmp/one.log 2>&1
  r=$(grep '"pass"' /tmp/one.log | tail -8 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=[x['reads_per_s'] for x in v]
print(int(sum(r)/n), int(min(r)), int(max(r)), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2), 'dp_ms', round(1e3*sum(x['dp_wall_s'] for x in v)/n,2), 'seed_ms', round(1e3*sum(x['seed_wall_s'] for x in v)/n,2), 'upl_ms', round(1e3*sum(x['upload_s'] for x in v)/n,2), 'rounds', v[-1]['dp_rounds'], 'parse_ms', round(1e3*sum(x['read_parse_s'] for x in v)/n,2))")
  echo "$lab : $r"; grep "ya_sw_batch wall" /tmp/one.log
}
export YA_SYNC=spin YA_PROF=1 YA_NO_GPU_LOCK=1
for pool in 16 13; do
for cfg in "2500 4 0" "2500 8 0" "1250 8 0" "1250 16 0" "2500 4 100" "5000 4 0" "1000 20 0"; do
  set -- $cfg
  YA_COALESCE_US=$3 one "nolock e2e pool=$pool batch=$1 pipes=$2 co=$3" -batch $1 -pipes $2 -tpp $pool -passes 12
done; done
YA_SYNC=block YA_COALESCE_US=0 one "nolock BLOCK e2e pool=16 batch=2500 pipes=8" -batch 2500 -pipes 8 -passes 12
YA_SYNC=block YA_COALESCE_US=0 one "nolock BLOCK e2e pool=16 batch=1250 pipes=16" -batch 1250 -pipes 16 -passes 12
