#!/usr/bin/env python
"""Does polling NVML disturb a running job?  Runs the cfg3 e2e job (-passes 60) with and without a 20 Hz NVML poll thread and
prints the step times and the duration of every NVML query.  (bench.py samples clocks during its timed region.)"""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = "/tmp/yaha_b200_bench_iid100"
X = [os.path.join(D, f) for f in os.listdir(D) if f.startswith("ref.X15_01_")][0]
cmd = [os.path.join(ROOT, "yaha_b200", "yaha_b200_host"), "-x", X, "-q", D + "/reads_rank0.fa", "-osh", "/tmp/o_j.sam", "-t", "4", "-batch", "5000",
       "-pipes", "4", "-passes", "60", "-BW", "10", "-G", "100"]
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for mode in ("none", "clock_only", "clock+reasons"):
    stop = [False]; durs = []
    def poll():
        while not stop[0]:
            t = time.perf_counter()
            if mode != "none":
                pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                if mode == "clock+reasons":
                    pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            durs.append((time.perf_counter() - t) * 1e3)
            time.sleep(0.05)
    th = threading.Thread(target=poll); th.start()
    p = subprocess.run(cmd, capture_output=True, text=True)
    stop[0] = True; th.join()
    st = [json.loads(l)["align_s"] * 1e3 for l in p.stderr.splitlines() if l.startswith('{"pass"')][8:]
    print(mode, "steps ms: median %.2f max %.2f  >12ms: %d of %d" % (sorted(st)[len(st) // 2], max(st), sum(1 for x in st if x > 12), len(st)),
          "| nvml query ms: max %.2f mean %.2f n=%d" % (max(durs), sum(durs) / len(durs), len(durs)))
