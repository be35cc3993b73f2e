#!/usr/bin/env python
"""bench.py -- reads/s and banded-SW GCUPS of the yaha alignment hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg3|cfg2|cfg1s]

One "step" = one pass of the hot path (seed lookup -> hits->fragments->regions -> banded
affine-gap DP with X-drop + traceback) over the whole synthetic read set of the workload.
Workloads follow BASELINE.json `configs`:
  cfg3 (default): 100 Mbp i.i.d. reference, 20 000 x 500 bp reads at 10 % error, -BW 10 -G 100
                  (SW-extension-bound; the config the metric quotes at 1/2/4/8 B200)
  cfg2          : 100 Mbp reference, 100 000 x 100 bp reads at 5 % error (seed-lookup-bound)
  cfg1s         : 10 Mbp reference, 10 000 x 1000 bp reads at 2 % (the CPU-runnable case)
Multi-GPU: one process per GPU (torchrun), replicated index, every rank aligns its own read set
of the same shape (weak scaling), no collective on the data path; time = max over ranks.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/yaha, built by oracle/Makefile)
on the host cores with `-t <all cores>` on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ref_bases, n_reads, read_len, err, flags)
    "cfg3": (100_000_000, 20_000, 500, 0.10, dict(bw=10, max_gap=100)),
    "cfg2": (100_000_000, 100_000, 100, 0.05, dict()),
    "cfg1s": (10_000_000, 10_000, 1000, 0.02, dict()),
}
REF_FLAGS = {"cfg3": ["-BW", "10", "-G", "100"], "cfg2": [], "cfg1s": []}
WORKLOAD_DESC = {
    "cfg3": "BASELINE configs[2]: synthetic 100 Mbp reference, 20K 500 bp reads at 10% error, -BW 10 -G 100",
    "cfg2": "BASELINE configs[1]: synthetic 100 Mbp reference, 100K 100 bp reads at 5% error",
    "cfg1s": "BASELINE configs[0]: synthetic 10 Mbp reference, 10K 1000 bp reads at 2% error",
}
INT_OPS_PER_CELL_EXT = 33      # algorithmic integer ops per extension cell


def cache_dir(workload: str) -> str:
    d = os.path.join(os.environ.get("YAHA_BENCH_CACHE", tempfile.gettempdir()), f"yaha_b200_bench_{workload}")
    os.makedirs(d, exist_ok=True)
    return d


def make_reference(workload: str):
    """Reference FASTA + .nib2 on disk (cached); returns (dir, Nib2)."""
    from yaha_b200 import refio, synth
    nb = WORKLOADS[workload][0]
    d = cache_dir(workload)
    fa, nib = os.path.join(d, "ref.fa"), os.path.join(d, "ref.nib2")
    if not os.path.exists(nib):
        ref = synth.random_reference(nb, 12345)
        synth.write_fasta(fa + ".tmp", [("chr1", ref)])
        os.replace(fa + ".tmp", fa)
        img = refio.build_nib2([("chr1", ref)])
        with open(nib + ".tmp", "wb") as f:
            f.write(img)
        os.replace(nib + ".tmp", nib)
    return d, refio.load_nib2(nib)


def make_reads(workload: str, rank: int):
    """(names, list of char arrays) for this rank's shard; deterministic per rank."""
    from yaha_b200 import synth
    nb, n_reads, rl, err, _ = WORKLOADS[workload]
    ref = synth.random_reference(nb, 12345)
    reads = list(synth.simulate_reads(ref, n_reads, rl, err, 777 + 1000 * rank))
    return reads


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "samples": len(self.rows)}


def extension_jobs_from_frags(strands, frags, n_reads, read_lens, max_roff, JOB_DT):
    """Host-side caller stand-in until the batched host driver lands: for every read take the
    longest surviving seed fragment of either strand and issue the backward + forward X-drop
    extensions extendClumpForwardReverse would (AlignExtFrag.cpp:64-141, without the perfect
    pre-extension)."""
    nfr = strands["n_frags"].astype(np.int64)
    seg = np.repeat(np.arange(2 * n_reads, dtype=np.int64), nfr)
    if len(seg) == 0:
        return np.zeros(0, dtype=JOB_DT)
    rd = seg >> 1
    order = np.lexsort((-frags["refLen"].astype(np.int64), rd))
    rds, firsts = np.unique(rd[order], return_index=True)
    k = order[firsts]
    st = (seg[k] & 1).astype(np.uint8)
    f = frags[k]
    sqo = f["startQueryOff"].astype(np.int64); eqo = f["endQueryOff"].astype(np.int64)
    sro = f["startRefOff"].astype(np.int64); ero = sro + f["refLen"].astype(np.int64) - 1
    back = np.minimum(sqo, sro)
    forw = np.minimum(read_lens[rds] - 1 - eqo, max_roff - ero)
    jb = np.zeros(len(k), dtype=JOB_DT)
    jb["rOff"] = sro - 1; jb["read"] = rds; jb["qOff"] = sqo - 1; jb["qLen"] = back; jb["kind"] = 3; jb["strand"] = st
    jf = np.zeros(len(k), dtype=JOB_DT)
    jf["rOff"] = ero + 1; jf["read"] = rds; jf["qOff"] = eqo + 1; jf["qLen"] = forw; jf["kind"] = 2; jf["strand"] = st
    return np.concatenate([jb[back >= 5], jf[forw >= 5]])


def run_ours(args):
    import torch
    import torch.distributed as dist
    import yaha_b200
    from yaha_b200 import refio

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload
    nb, n_reads, rl, err, flags = WORKLOADS[wl]
    d, nib = make_reference(wl) if rank == 0 or world == 1 else (None, None)
    if world > 1:
        dist.barrier()
        if rank != 0:
            d, nib = make_reference(wl)
    params = yaha_b200.Params.defaults(word_len=15, **flags)
    t0 = time.time()
    al = yaha_b200.Aligner(nib, None, params, device=local)         # index built on the device
    t_index = time.time() - t0

    reads = make_reads(wl, rank)
    fwd = [refio.encode(s) for _, s in reads]
    lens = np.array([len(c) for c in fwd], dtype=np.int64)
    offs = np.zeros(len(fwd) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    codes_host = torch.from_numpy(np.concatenate(fwd)).pin_memory()
    codes_np = codes_host.numpy()
    h2d_bytes = int(codes_np.nbytes + offs.nbytes)

    def step(upload: bool):
        if upload:
            al.upload_reads(codes_np, offs)
        strands, frags, region = al.seed_frags(frags_cap=1 << 22)
        jobs = extension_jobs_from_frags(strands, frags, len(fwd), lens, nib.max_roff, yaha_b200.JOB_DT)
        res, ops = al.sw_batch(jobs, ops_cap=1 << 24)
        return strands, frags, jobs, res, ops

    al.upload_reads(codes_np, offs)
    for _ in range(args.warmup):
        out = step(True)
    d2h_bytes = int(out[0].nbytes + out[1].nbytes + len(out[1]) * 4 + out[3].nbytes + out[4].nbytes)
    al.counters()

    def timed(upload: bool):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            step(upload)
        torch.cuda.synchronize()
        el = time.perf_counter() - t
        if world > 1:
            tt = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.item())
        return el

    sampler = ClockSampler(local)
    sampler.start()
    el_resident = timed(False)
    ctr = al.counters()
    el_e2e = timed(True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    al.counters()

    total_reads = n_reads * world * args.steps
    value = total_reads / el_resident
    e2e = total_reads / el_e2e
    gcups = ctr.dp_cells / (ctr.ms_dp * 1e-3) / 1e9 if ctr.ms_dp > 0 else 0.0
    int_add, int_mix = al.int32_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650"
    # dominant kernel of this workload: the banded-SW fill (dp_wave_kernel); INT32-issue bound
    achieved_giops = gcups * INT_OPS_PER_CELL_EXT
    # seed stage: algorithmic bytes = 8 B per probe (SO[h],SO[h+1]) + 4 B per hit + one ideal sort pass
    # (16 B per hit) + 12 B per fragment, SURVEY.md section 8(d)
    seed_bytes = 8.0 * ctr.probes + 20.0 * ctr.hits + 12.0 * ctr.frags_all
    seed_gbs = seed_bytes / (ctr.ms_seed * 1e-3) / 1e9 if ctr.ms_seed > 0 else 0.0

    line = {
        "metric": "reads/s (hot path: seed lookup + hits->fragments + banded-SW extension) and banded-SW GCUPS",
        "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el_resident / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[wl], "reads_per_gpu": n_reads, "read_len": rl, "error": err,
                   "flags": REF_FLAGS[wl], "l2": "inputs larger than L2 (4.3 GB index + reads re-read per step)",
                   "dp_jobs": "2 X-drop extensions per read from its longest seed fragment (host graph/split stages not in the timed path yet)",
                   "index_build_s": round(t_index, 2)},
        "gcups": gcups, "dp_cells_per_step": ctr.dp_cells // max(args.steps, 1),
        "stage_ms_per_step": {"seed": ctr.ms_seed / args.steps, "dp_fill": ctr.ms_dp / args.steps,
                              "traceback": ctr.ms_traceback / args.steps},
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": int(ctr.launches),
        "roofline": {"bound": "int32-issue", "kernel": "dp_wave_kernel", "achieved": achieved_giops, "peak": int_mix,
                     "unit": "GIOP/s", "frac": achieved_giops / int_mix if int_mix else None, "traffic": None,
                     "peak_source": "ya_measure_int32_peak (DP-cell op mix, measured live)",
                     "peak_add_only": int_add, "ops_per_cell": INT_OPS_PER_CELL_EXT},
        "roofline_seed": {"bound": "hbm", "kernel": "seed_count + expand + radix + frag scans", "achieved": seed_gbs,
                          "peak": hbm_peak, "unit": "GB/s", "frac": seed_gbs / hbm_peak, "traffic": None,
                          "peak_source": hbm_src},
        "clocks": sampler.summary(),
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, d, nib, al, reads, sample=min(n_reads, args.cpu_sample))
        print(json.dumps(line), flush=True)
    al.close()
    if world > 1:
        dist.destroy_process_group()


def ensure_index_file(wl: str, d: str, al=None) -> str:
    """Index file on disk for the reference binary.  Built by the reference itself (`yaha -g`)
    unless an Aligner with a device-built (bit-identical) index is at hand."""
    from yaha_b200 import refio
    path = os.path.join(d, refio.index_file_name("ref", 15, 1, 65525))
    if os.path.exists(path):
        return path
    if al is not None:
        idx = al.download_index()
        refio.write_index(path + ".tmp", idx)
        os.replace(path + ".tmp", path)
    else:
        yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
        subprocess.check_call([yaha, "-g", "ref.fa", "-L", "15", "-S", "1"], cwd=d,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return path


def time_reference(wl: str, d: str, index_path: str, reads, threads: int, runs: int = 1, warm: int = 1):
    from yaha_b200 import synth
    yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
    if not os.path.exists(yaha):
        return None
    qf = os.path.join(d, f"sample_{len(reads)}.fa")
    synth.write_reads(qf, reads)
    cmd = [yaha, "-x", index_path, "-q", qf, "-osh", os.path.join(d, "ref_out.sam"), "-t", str(threads)] + REF_FLAGS[wl]
    best = None
    for k in range(warm + runs):
        t = time.perf_counter()
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        el = time.perf_counter() - t
        if k >= warm:
            best = el if best is None else min(best, el)
    return best


def cpu_baseline(wl, d, nib, al, reads, sample):
    ncores = os.cpu_count() or 1
    idx = ensure_index_file(wl, d, al)
    sub = reads[:sample]
    el = time_reference(wl, d, idx, sub, ncores)
    if el is None:
        return {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "oracle/_ref/yaha missing"}
    return {"value": len(sub) / el, "unit": "reads/s", "cores": ncores, "kind": "reference",
            "sample": f"{len(sub)} reads of the workload, whole program wall time (index mmap + page-touch included), "
                      f"yaha -t {ncores}, best of 1 after 1 warm-up pass"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    nb, n_reads, rl, err, flags = WORKLOADS[wl]
    yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
    if not os.path.exists(yaha):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/yaha not built (reference sources absent)"}))
        return
    d, nib = make_reference(wl)
    idx = ensure_index_file(wl, d, None)
    reads = make_reads(wl, 0)
    sample = reads[:min(n_reads, args.cpu_sample)]
    ncores = os.cpu_count() or 1
    times = []
    for k in range(args.warmup + args.steps):
        el = time_reference(wl, d, idx, sample, ncores, runs=1, warm=0)
        if k >= args.warmup:
            times.append(el)
    tot = sum(times)
    v = len(sample) * len(times) / tot
    line = {"impl": "reference", "metric": "reads/s (whole yaha program, all stages, SAM written)", "value": v, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[wl], "flags": REF_FLAGS[wl], "threads": ncores},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": ncores, "kind": "reference",
                             "sample": f"{len(sample)} reads per step, yaha -t {ncores}, whole program wall time"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
