#!/bin/bash
# Run on the GPU box.  Captures ONE launch of each kernel named in $KERNELS with `ncu --set full` (whole-shard batch: the
# launch sizes of bench.py's run C) plus, for the extension kernel, one launch of the value-run size (5000-read batch);
# raw CSV pages go to gpurun_out/ (summarised into profiles/ by tools/ncu_summary.py).
set -u
python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
D=${YAHA_BENCH_CACHE:-/tmp}/yaha_b200_bench_iid100
X=$(ls $D/ref.X15_01_* | head -1); Q=$D/reads_rank0.fa
H=yaha_b200/yaha_b200_host; O=gpurun_out; mkdir -p $O
RUN="$H -x $X -q $Q -osh /tmp/ncu_o.sam -t 4 -batch 20000 -pipes 1 -BW 10 -G 100"
for k in ${KERNELS:-dp_ext_packed seed_count traceback_warp k2_fused form_clumps assemble_kernel finish_kernel format_kernel}; do
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/prof_$k $RUN > $O/ncu_$k.log 2>&1
  ncu -i $O/prof_$k.ncu-rep --page raw --csv > $O/prof_$k.raw.csv 2>/dev/null
  rm -f $O/prof_$k.ncu-rep
done
RUN5="$H -x $X -q $Q -osh /tmp/ncu_o.sam -t 4 -batch 5000 -pipes 1 -BW 10 -G 100"
timeout -s KILL 600 ncu --set full --clock-control none -k regex:dp_ext_packed -c 1 -f -o $O/prof_dp_ext_5000 $RUN5 > $O/ncu_dp_ext_5000.log 2>&1
ncu -i $O/prof_dp_ext_5000.ncu-rep --page raw --csv > $O/prof_dp_ext_5000.raw.csv 2>/dev/null
rm -f $O/prof_dp_ext_5000.ncu-rep
echo done
