#!/bin/bash
# value-run configurations (replayed reads): large lock-step batches vs many small ones
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { # label, extra host args...
  lab=$1; shift
  yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 "$@" > /tmp/one.log 2>&1
  r=$(grep '"pass"' /tmp/one.log | tail -12 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=sorted(x['reads_per_s'] for x in v)
ec=sum(x['ext_cells'] for x in v); em=sum(x['dev_ms_ext'] for x in v); el=sum(x['ext_launches'] for x in v)
print(int(sum(r)/n), 'med', int(r[n//2]), int(min(r)), int(max(r)), 'rounds', v[-1]['dp_rounds'], 'ext_gcups', round(ec/max(em,1e-9)/1e6,1), 'ext_launches/pass', el/n, 'ms/launch', round(em/max(el,1),3))")
  echo "$lab : $r"
}
for lock in 0 1; do
for co in 0 500 3000; do
for cfg in "10000 2" "5000 4" "20000 1"; do
  set -- $cfg
  if [ $lock = 1 ]; then export YA_GPU_LOCK=1; else unset YA_GPU_LOCK; fi
  YA_COALESCE_US=$co one "replay lock=$lock co=$co batch=$1 pipes=$2" -batch $1 -pipes $2 -passes 16 -replay
done; done; done
