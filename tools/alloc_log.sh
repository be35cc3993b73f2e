#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
YA_ALLOC_LOG=1 yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -BW 10 -G 100 -batch 2500 -pipes 8 -passes 20 2> /tmp/al.log
python - <<PY
import json,re
t_pass=[]; allocs=[]
for l in open('/tmp/al.log'):
    if l.startswith('{"pass"'):
        d=json.loads(l); t_pass.append((d['pass'], int(d['reads_per_s'])))
        print('pass', d['pass'], int(d['reads_per_s']), 'allocs since last:', len(allocs), 'ms', round(sum(a for a in allocs),2)); allocs=[]
    elif l.startswith('ya_alloc'):
        m=re.search(r'([\d.]+) ms', l); allocs.append(float(m.group(1)))
        if len(t_pass)>=4 and 'host_free' not in l: print('   late:', l.strip())
PY
