#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { # label, extra host args...
  lab=$1; shift
  yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 "$@" > /tmp/one.log 2>&1
  r=$(grep '"pass"' /tmp/one.log | tail -12 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=sorted(x['reads_per_s'] for x in v)
print(int(sum(r)/n), 'med', int(r[n//2]), int(min(r)), int(max(r)), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2), 'dp_ms', round(1e3*sum(x['dp_wall_s'] for x in v)/n,2), 'seed_ms', round(1e3*sum(x['seed_wall_s'] for x in v)/n,2), 'rounds', v[-1]['dp_rounds'], 'parse_ms', round(1e3*sum(x['read_parse_s'] for x in v)/n,2))")
  echo "$lab : $r"
}
export YA_COALESCE_US=0 YA_NO_GPU_LOCK=1
for sync in nap yield spin; do
for pool in 16 14; do
for cfg in "2500 4" "2500 8" "1250 8"; do
  set -- $cfg
  YA_SYNC=$sync one "e2e sync=$sync pool=$pool batch=$1 pipes=$2" -batch $1 -pipes $2 -tpp $pool -passes 20
done; done; done
for cfg in "2500 8" "5000 4" "1250 16"; do set -- $cfg; one "replay nap batch=$1 pipes=$2" -batch $1 -pipes $2 -passes 20 -replay; done
YA_TRACE=gpurun_out/trace_d.txt yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 -batch 2500 -pipes 8 -passes 12 2>&1 | grep '"pass"' | tail -1 | cut -c1-120
python tools/trace_view.py gpurun_out/trace_d.txt
