#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
MODE=${1:-}
YA_TRACE=gpurun_out/trace_e.txt yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 -batch 2500 -pipes 8 -passes 14 $MODE 2>&1 | grep '"pass"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['pass'], int(d['reads_per_s']), 'rounds', d['dp_rounds'], 'parse', d['read_parse_s'], 'write', d['write_s'])"
python tools/trace_view.py gpurun_out/trace_e.txt 12
python tools/trace_view.py gpurun_out/trace_e.txt 13
