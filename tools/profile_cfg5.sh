#!/bin/bash
# Kernel launch list (durations) of one 5000-read batch of the cfg5 workload (3.1 Gbp reference, Alu-like family, -H 650 -MD 50).
export YAHA_BENCH_CACHE=${YAHA_BENCH_CACHE:-/tmp/ybc}; mkdir -p $YAHA_BENCH_CACHE
YAHA_BENCH_CFG5_READS=5000 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/cfg5_small.json 2> gpurun_out/cfg5_small.err
D=$(ls -d $YAHA_BENCH_CACHE/yaha_b200_bench_human*); X=$(ls $D/ref.X15_01_* | head -1); Q=$D/reads_rank0.fa
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg5.csv \
    yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/c5.sam -t 8 -batch 5000 -pipes 1 -H 650 -MD 50 > gpurun_out/ncu_cfg5.log 2>&1
python tools/launch_summary.py gpurun_out/launches_cfg5.csv > gpurun_out/launches_cfg5.md
head -30 gpurun_out/launches_cfg5.md
YAHA_B200_STATS=1 yaha_b200/yaha_b200_host -x $X -q $Q -osh /tmp/c5.sam -t 8 -batch 5000 -pipes 1 -passes 3 -H 650 -MD 50 2>&1 | grep '"pass"' | tail -1 | python3 -c "
import sys, json
d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('align_s','dev_ms_seed','dev_ms_lookup','dev_ms_dp','dev_ms_traceback','dev_ms_finish','hits','probes','frags_all','reads_handed_back','dp_jobs','launches')})"
