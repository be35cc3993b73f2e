#!/bin/bash
D=/tmp/yaha_b200_bench_cfg1s
python bench.py --workload cfg1s --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
YA_TRACE=gpurun_out/trace_1s.txt YA_ALLOC_LOG=1 yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -batch 2500 -pipes 8 -passes 8 2> /tmp/e.log
grep '"pass"' /tmp/e.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['pass'], int(d['reads_per_s']), 'rounds', d['dp_rounds'], 'parse', d['read_parse_s'], 'write', d['write_s'], 'upload', d['upload_s'], 'seed', d['seed_wall_s'], 'dp', d['dp_wall_s'], 'host', d['host_wall_s'])"
grep -c ya_alloc /tmp/e.log; grep ya_alloc /tmp/e.log | grep -v host_free | tail -5
python tools/trace_view.py gpurun_out/trace_1s.txt 7
