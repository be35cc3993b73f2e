#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
for rep in 1 2; do
for cfg in "10000 2 8" "5000 4 8" "5000 4 4" "2500 4 8" "2500 8 4" "5000 3 8" "4000 5 8" "2500 8 8" "10000 2 16"; do
  set -- $cfg
  r=$(yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -batch $1 -pipes $2 -tpp $3 -passes 16 -replay -BW 10 -G 100 2>&1 | grep '"pass"' | tail -12 | python -c "
import sys,json
v=[json.loads(l)['reads_per_s'] for l in sys.stdin]; print(int(sum(v)/len(v)), int(min(v)), int(max(v)))")
  echo "batch=$1 pipes=$2 tpp=$3 : $r"
done; done
