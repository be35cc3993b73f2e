#!/bin/bash
# Run on the GPU box after bench.py has filled its cache: kernel launch list (ncu, durations only) of the fused path for one
# pipeline, and a host timeline of an e2e pass.  usage: tools/profile_fused.sh [batch] [tag] [pipes of the launch list]
B=${1:-5000}; TAG=${2:-v1}; NP=${3:-1}
D=/tmp/yaha_b200_bench_iid100; X=$(ls $D/ref.X15_01_* | head -1); Q=$D/reads_rank0.fa
H=yaha_b200/yaha_b200_host; O=gpurun_out; mkdir -p $O
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_fused_$TAG.csv \
    $H -x $X -q $Q -osh /tmp/ncu_o.sam -t 4 -batch $B -pipes $NP -BW 10 -G 100 -passes 2 > $O/ncu_list_$TAG.log 2>&1
python tools/launch_summary.py $O/launches_fused_$TAG.csv > $O/launches_fused_$TAG.md
cat $O/launches_fused_$TAG.md | head -40
YA_TRACE=$O/trace_fused_$TAG.txt $H -x $X -q $Q -osh /tmp/sweep.sam -t 4 -BW 10 -G 100 -batch $B -pipes 4 -passes 12 2>&1 | grep '"pass"' | python -c "
import sys,json
print('e2e ms per pass', [round(json.loads(l)['align_s']*1e3,1) for l in sys.stdin])"
python tools/trace_view.py $O/trace_fused_$TAG.txt 10 | head -60
