#!/bin/bash
D=/tmp/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
for p in 4 24; do
YA_PROF=1 yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/sweep.sam -t 16 -batch 10000 -pipes 2 -passes $p -replay -BW 10 -G 100 2>&1 | grep "ya_sw_batch wall"
done
