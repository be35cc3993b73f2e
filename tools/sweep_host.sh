#!/bin/bash
# A/B of host binaries / allocator settings on the GPU box (replayed passes: no parse, no write)
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
run() { # name binary
  r=$($2 -x $X -q $Q -osh /tmp/sweep.sam -t 16 -batch 10000 -pipes 2 -passes 16 -replay -BW 10 -G 100 2>&1 | grep '"pass"' | tail -12 | python -c "
import sys,json
v=[json.loads(l) for l in sys.stdin]; n=len(v); print(int(sum(x['reads_per_s'] for x in v)/n), 'host_ms', round(1e3*sum(x['host_wall_s'] for x in v)/n,2), 'dp_ms', round(1e3*sum(x['dp_wall_s'] for x in v)/n,2), 'seed_ms', round(1e3*sum(x['seed_wall_s'] for x in v)/n,2), 'upl_ms', round(1e3*sum(x['upload_s'] for x in v)/n,2))")
  echo "$1 : $r"
}
for rep in 1 2 3; do
  run prev yaha_b200/yaha_b200_host_prev
  run new yaha_b200/yaha_b200_host
  GLIBC_TUNABLES=glibc.malloc.tcache_count=127 run new_tcache127 yaha_b200/yaha_b200_host
  GLIBC_TUNABLES=glibc.malloc.tcache_count=127:glibc.malloc.mxfast=160 run new_tc127_mxfast yaha_b200/yaha_b200_host
done
