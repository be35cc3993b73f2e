#!/bin/bash
# e2e sweep over batch size x pipelines (cfg3 job, 16 worker threads)
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { lab=$1; shift
  yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 "$@" > /tmp/one.log 2>&1
  grep '"pass"' /tmp/one.log | tail -30 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=sorted(x['align_s']*1e3 for x in v)
print('$lab', 'median ms', round(r[n//2],2), 'mean', round(sum(r)/n,2), 'min', round(r[0],2), 'max', round(r[-1],2), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2), 'rounds', v[-1]['dp_rounds'])"; }
for cfg in ${CFGS:-"1250 8" "1250 10" "1250 12" "1250 16" "1000 10" "1000 12" "834 12" "2500 8"}; do set -- $cfg
  one "e2e batch=$1 pipes=$2" -batch $1 -pipes $2 -passes 40
done
