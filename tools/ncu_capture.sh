#!/bin/bash
# Run on the GPU box.  Captures one launch of the kernels named in $KERNELS with `ncu --set full` plus the launch
# list of a whole cfg3 job (one 20 K-read batch, lock-step rounds); outputs go to gpurun_out/ (summarised into profiles/).
set -u
D=${YAHA_BENCH_CACHE:-/tmp}/yaha_b200_bench_cfg3
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
X=$D/ref.X15_01_65525S
Q=$D/reads_rank0.fa
H=yaha_b200/yaha_b200_host
O=gpurun_out
mkdir -p $O
export YA_COALESCE_US=20000
RUN="$H -x $X -q $Q -osh /tmp/ncu_o.sam -t 16 -batch 20000 -pipes 1 -BW 10 -G 100"
for k in ${KERNELS:-traceback_warp}; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/prof_$k $RUN > $O/ncu_$k.log 2>&1
  ncu -i $O/prof_$k.ncu-rep --page raw --csv > $O/prof_$k.raw.csv 2>/dev/null
done
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_v3.csv $RUN -passes 2 > $O/ncu_list.log 2>&1
echo done
