#!/bin/bash
# End-of-round calls on the GPU box.  part 1: GPU suite, smoke, differential fuzzing of the CUDA path against oracle/_ref/yaha.
# part 2: the value run's launch list, one `ncu --set full` capture of its extension launch (10 000-read batch) and the cfg5 line.
export YAHA_BENCH_CACHE=/tmp/ybc; mkdir -p $YAHA_BENCH_CACHE gpurun_out
O=gpurun_out; H=yaha_b200/yaha_b200_host
if [ "${1:-1}" = 1 ]; then
  python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
  python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
  timeout 150 python tools/fuzz_parity.py --seeds 20000:20100 --jobs 8 --binary $H --keep $O/fuzz_fail --wordlens 11,11,12,13 > $O/fuzz_gpu.log 2>&1
  tail -1 $O/fuzz_gpu.log; grep -c '^ok' $O/fuzz_gpu.log; grep -A3 FAIL $O/fuzz_gpu.log | head -20
  timeout 120 python tools/fuzz_parity.py --heavy --seeds 21000:21024 --jobs 8 --binary $H --keep $O/fuzz_fail --wordlens 11,12,13 > $O/fuzz_gpu_heavy.log 2>&1
  tail -1 $O/fuzz_gpu_heavy.log; grep -c '^ok' $O/fuzz_gpu_heavy.log; grep -A3 FAIL $O/fuzz_gpu_heavy.log | head -20
else
  python bench.py --no-cpu-baseline --steps 3 --warmup 3 > /dev/null 2>&1
  D=$YAHA_BENCH_CACHE/yaha_b200_bench_iid100; X=$(ls $D/ref.X15_01_* | head -1); Q=$D/reads_rank0.fa
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_value_v12.csv \
      $H -x $X -q $Q -osh /tmp/ncu_o.sam -t 4 -batch 10000 -pipes 2 -BW 10 -G 100 -passes 2 > $O/ncu_list_v12.log 2>&1
  python tools/launch_summary.py $O/launches_value_v12.csv > $O/launches_value_v12.md; head -12 $O/launches_value_v12.md
  timeout -s KILL 300 ncu --set full --clock-control none -k regex:dp_ext_packed -c 1 -f -o $O/prof_dp_ext_10000 \
      $H -x $X -q $Q -osh /tmp/ncu_o.sam -t 4 -batch 10000 -pipes 1 -BW 10 -G 100 > $O/ncu_dp_ext_10000.log 2>&1
  ncu -i $O/prof_dp_ext_10000.ncu-rep --page raw --csv > $O/prof_dp_ext_10000.raw.csv 2>/dev/null; rm -f $O/prof_dp_ext_10000.ncu-rep
  timeout 400 python bench.py --workload cfg5 --batch 5000 --pipes 4 --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg5_v3.json 2> $O/r02_bench_cfg5_v3.err
  python - <<PY
import json
d=json.loads(open("$O/r02_bench_cfg5_v3.json").read().strip().splitlines()[-1])
print("cfg5", d["value"], d["e2e"]["value"], d["ms_per_step"], d["stage_ms_per_step"]["device_seed"])
PY
fi
