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
export YA_SYNC=spin YA_PROF=1
cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc
for pool in 16 14 12 10; do
for cfg in "2500 4 0" "2500 2 0" "5000 2 100" "5000 4 0"; do
  set -- $cfg
  YA_COALESCE_US=$3 one "e2e pool=$pool batch=$1 pipes=$2 co=$3" -batch $1 -pipes $2 -tpp $pool -passes 12
done; done
