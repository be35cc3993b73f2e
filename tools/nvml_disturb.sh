#!/bin/bash
# Does polling NVML while the host runs cause slow passes?  (cfg1s, 60 passes, with and without a 50 ms poller)
D=/tmp/yaha_b200_bench_cfg1s
python bench.py --workload cfg1s --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
X=$D/ref.X15_01_65525S; Q=$D/reads_rank0.fa
run() { yaha_b200/yaha_b200_host -x $X -q $Q -osh $D/out_x.sam -t 16 -batch 2500 -pipes 8 -passes 60 2>&1 | grep '"pass"' | python -c "
import sys,json
v=[round(json.loads(l)['align_s']*1e3,1) for l in sys.stdin][5:]
print('$1', 'max', max(v), 'median', sorted(v)[len(v)//2], 'slow(>25ms):', [x for x in v if x>25])"; }
run none
python - <<PY &
import pynvml, time
pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
t=time.time()
while time.time()-t<12:
    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h); time.sleep(0.05)
PY
sleep 1; run nvml50ms; wait
python - <<PY &
import pynvml, time
pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
t=time.time()
while time.time()-t<12:
    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); time.sleep(0.05)
PY
sleep 1; run nvml_clock_only; wait
run none_again
