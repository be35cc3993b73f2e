#!/bin/bash
# Timeline of one e2e (or -replay) pass of the cfg3 job: tools/trace_run.sh [-replay]
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
MODE=${1:-}
YA_TRACE=gpurun_out/trace_e.txt yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 -batch ${BATCH:-1250} -pipes 8 -passes 16 $MODE 2>&1 | grep '"pass"' | python -c "
import sys,json
print([round(json.loads(l)['align_s']*1e3,1) for l in sys.stdin])"
python tools/trace_view.py gpurun_out/trace_e.txt 14
