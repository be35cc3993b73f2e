#!/bin/bash
# Run on the GPU box after `python bench.py` has populated the workload cache (/tmp/yaha_b200_bench_cfg3).
# Captures one launch each of the two dominant kernels with `ncu --set full` plus the launch list of a
# whole cfg3 job; outputs go to gpurun_out/ (summarised by hand into profiles/).
set -u
D=${YAHA_BENCH_CACHE:-/tmp}/yaha_b200_bench_cfg3
X=$D/ref.X15_01_65525S
Q=$D/reads_rank0.fa
H=yaha_b200/yaha_b200_host
O=gpurun_out
mkdir -p $O
RUN="$H -x $X -q $Q -osh /tmp/ncu_o.sam -t 16 -batch 20000 -pipes 1 -BW 10 -G 100"
YAHA_B200_STATS=1 $RUN 2> $O/ncu_plain_stats.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:dp_ext_packed -c 1 -f -o $O/prof_ext_packed $RUN > $O/ncu_ext.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:seed_count -c 1 -f -o $O/prof_seed_count $RUN > $O/ncu_seed.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:traceback_kernel -s 1 -c 1 -f -o $O/prof_traceback $RUN > $O/ncu_tb.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_final.csv $RUN > $O/ncu_list.log 2>&1
echo done
