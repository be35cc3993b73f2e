#!/bin/bash
# e2e / replay sweep of the shared-worker-pool host on the GPU box: batch x pipelines x coalescing wait x sync mode
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { # label, extra host args...
  lab=$1; shift
  r=$(yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 "$@" 2>&1 | grep '"pass"' | tail -8 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=[x['reads_per_s'] for x in v]
print(int(sum(r)/n), int(min(r)), int(max(r)), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2), 'dp_ms', round(1e3*sum(x['dp_wall_s'] for x in v)/n,2), 'rounds', v[-1]['dp_rounds'], 'parse_ms', round(1e3*sum(x['read_parse_s'] for x in v)/n,2), 'write_ms', round(1e3*sum(x['write_s'] for x in v)/n,2))")
  echo "$lab : $r"
}
for sync in block spin; do
for co in 0 100 300; do
for cfg in "2500 4" "2500 8" "1250 8" "5000 4" "5000 2" "1000 12"; do
  set -- $cfg
  YA_SYNC=$sync YA_COALESCE_US=$co one "e2e sync=$sync co=$co batch=$1 pipes=$2" -batch $1 -pipes $2 -passes 12
done; done; done
for co in 0 100 300; do
for cfg in "2500 4" "2500 8" "5000 4" "10000 2"; do
  set -- $cfg
  YA_COALESCE_US=$co one "replay co=$co batch=$1 pipes=$2" -batch $1 -pipes $2 -passes 12 -replay
done; done
for tpp in 24 32; do YA_COALESCE_US=100 one "e2e pool=$tpp batch=2500 pipes=8" -batch 2500 -pipes 8 -tpp $tpp -passes 12; done
