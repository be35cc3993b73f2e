#!/bin/bash
# Throughput of a rank that owns only 4 or 8 cores (taskset confines the whole host process), with the fragment graph
# and phase 1 of the alignment on the device (default) or in the worker threads (YA_HOST_CLUMPS=1).
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
one() { lab=$1; cpus=$2; shift; shift
  taskset -c $cpus yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -BW 10 -G 100 "$@" > /tmp/one.log 2>&1
  grep '"pass"' /tmp/one.log | tail -12 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); r=sorted(x['align_s']*1e3 for x in v)
print('$lab', 'median ms', round(r[n//2],2), 'reads/s', int(20000/(sum(r)/n)*1e3), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2))"; }
for mode in dev host; do
  if [ $mode = host ]; then export YA_HOST_CLUMPS=1; else unset YA_HOST_CLUMPS; fi
  one "$mode 4 cores" 0-3 -t 4 -pipes 4 -batch 1250 -passes 16
  one "$mode 8 cores" 0-7 -t 8 -pipes 8 -batch 1250 -passes 16
done
