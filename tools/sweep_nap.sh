#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { lab=$1; shift
  yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 "$@" > /tmp/one.log 2>&1
  grep '"pass"' /tmp/one.log | tail -30 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=sorted(x['align_s']*1e3 for x in v)
print('$lab', 'median ms', round(r[n//2],2), 'mean', round(sum(r)/n,2), 'min', round(r[0],2), 'max', round(r[-1],2), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2), 'rounds', v[-1]['dp_rounds'])"; }
for rep in 1 2; do
YA_SYNC=callback one "e2e callback" -batch 1250 -pipes 8 -passes 40
YA_NAP_US=20 one "e2e nap=20" -batch 1250 -pipes 8 -passes 40
YA_NAP_US=40 one "e2e nap=40" -batch 1250 -pipes 8 -passes 40
YA_NAP_US=80 one "e2e nap=80" -batch 1250 -pipes 8 -passes 40
done
YA_SYNC=callback one "replay callback" -batch 2500 -pipes 8 -passes 40 -replay
YA_NAP_US=40 one "replay nap=40" -batch 2500 -pipes 8 -passes 40 -replay
