#!/bin/bash
# Find and show the slowest late pass of a long e2e run (timeline + allocation log).
W=${1:-cfg1s}
D=/tmp/yaha_b200_bench_$W
python bench.py --workload $W --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
FL=""; [ $W = cfg3 ] && FL="-BW 10 -G 100"
YA_TRACE=gpurun_out/trace_o.txt YA_ALLOC_LOG=1 yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -batch 2500 -pipes 8 -passes 120 $FL 2> /tmp/e.log
python - <<PY
import json, subprocess
ps=[json.loads(l) for l in open('/tmp/e.log') if l.startswith('{"pass"')]
late=[p for p in ps if p['pass']>=8]
late.sort(key=lambda p:-p['align_s'])
print('all ms:', [round(p['align_s']*1e3,1) for p in ps])
worst=late[0]['pass']
print('worst late pass', worst, late[0]['align_s'])
print(subprocess.run(['python','tools/trace_view.py','gpurun_out/trace_o.txt',str(worst)],capture_output=True,text=True).stdout)
PY
grep ya_alloc /tmp/e.log | grep -v host_free | tail -12
