#!/usr/bin/env python
"""bench.py -- reads/s and banded-SW GCUPS of the yaha alignment job on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg3|cfg1|cfg2|cfg4|cfg5]

One "step" = one complete alignment of the workload's synthetic read set by the product's host program
`yaha_b200/yaha_b200_host` (the call a user makes): the FASTA file is cut into records and parsed on the host, every batch of
reads goes to the device as characters, and ONE call of the C ABI (`ya_align_batch`) encodes them, runs the seed lookup,
hits -> fragments -> regions, the fragment graph, the first DP round with X-drop extensions and traceback (jobs born, laid out and
answered on the device), splices and scores every clump, runs OQC / filter by similarity and writes the SAM records; the host
writes the text.  The few reads whose clumps have to be split (0.16 % on cfg3) go through the call-by-call ABI in fibers.  The
SAM is byte-identical to the reference's (`cpu_baseline.sam_identical_to_reference`, `..._in_order_to_t1`).
`e2e` times all of that from the FASTA file to the SAM file; `value` replays the parsed reads from host memory and skips the
SAM fwrite (inputs resident; the H2D copy of the read characters is still inside).  Device stage times come from CUDA events
inside the library (`ya_get_counters`), on the stream the kernels are launched on.

Workloads follow BASELINE.json `configs`; cfg3 is the default (the config the metric quotes at
1/2/4/8 B200).  Multi-GPU: one process per GPU (torchrun), replicated index, every rank aligns its own read set of the same
shape (weak scaling), no collective on the data path; time = max over ranks.  Rank 0 additionally times the product's own
in-process path (`-gpus N`, one query file, `inprocess` in the JSON line).

`--impl reference` times the UNMODIFIED reference (oracle/_ref/yaha, built by oracle/Makefile) on the host cores with
`-t <all cores>` on the same workload.
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

HUMAN_BASES = int(float(os.environ.get("YAHA_BENCH_GBP", "3.1")) * 1e9)     # cfg4 / cfg5 reference size (scaled down on small boxes)
WORKLOADS = {
    # ref: "iid" = uniform ACGT, one sequence; "human" = 24 unequal sequences with an Alu-like repeat family and runs of N
    # reads: "sim" = uniform loci, per-base errors; "sv" = reads spanning implanted deletions / inversions / insertions
    # index_h: -H of the index build (650: over-full k-mers are down-sampled, Index.c:271-315)
    "cfg1": dict(ref="iid", ref_bases=10_000_000, reads="sim", n_reads=10_000, read_len=1000, err=0.02, flags=[], index_h=65525,
                 desc="BASELINE configs[0]: synthetic 10 Mbp reference, 10K 1000 bp reads at 2% error"),
    "cfg2": dict(ref="iid", ref_bases=100_000_000, reads="sim", n_reads=100_000, read_len=100, err=0.05, flags=[], index_h=65525,
                 desc="BASELINE configs[1]: synthetic 100 Mbp reference, 100K 100 bp reads at 5% error"),
    "cfg3": dict(ref="iid", ref_bases=100_000_000, reads="sim", n_reads=20_000, read_len=500, err=0.10, flags=["-BW", "10", "-G", "100"],
                 index_h=65525, desc="BASELINE configs[2]: synthetic 100 Mbp reference, 20K 500 bp reads at 10% error, -BW 10 -G 100"),
    "cfg4": dict(ref="human", ref_bases=HUMAN_BASES, reads="sv", n_reads=1_000, read_len=10_000, err=0.0, flags=["-OQC", "Y", "-FBS", "Y"],
                 index_h=650, desc=f"BASELINE configs[3]: synthetic {HUMAN_BASES / 1e9:.2f} Gbp reference in 24 sequences with an Alu-like family "
                                   "(one copy per 25 kbp) and N runs, index -H 650 (down-sampled), 1K 10 kbp reads spanning implanted "
                                   "deletions / inversions / insertions / translocated pieces at 2-5% error, -OQC Y -FBS Y"),
    "cfg5": dict(ref="human", ref_bases=HUMAN_BASES, reads="sim", n_reads=int(os.environ.get("YAHA_BENCH_CFG5_READS", "200000")), read_len=1000, err=0.05,
                 flags=["-H", "650", "-MD", "50"], index_h=650,
                 desc=f"BASELINE configs[4]: same {HUMAN_BASES / 1e9:.2f} Gbp reference, 1000 bp reads at 5% error, -H 650 -MD 50; a step is a bounded "
                      "sample of the 10M-read set"),
}
WORKLOADS["cfg1s"] = WORKLOADS["cfg1"]
REF_FLAGS = {k: v["flags"] for k, v in WORKLOADS.items()}
WORKLOAD_DESC = {k: v["desc"] for k, v in WORKLOADS.items()}
METRIC = "reads/s (whole alignment job: FASTA in -> SAM out, identical to reference)"
INT_OPS_PER_CELL_EXT = 33      # algorithmic integer ops per extension cell


def cache_dir(workload: str) -> str:
    w = WORKLOADS[workload]
    tag = "human%d" % (w["ref_bases"] // 1_000_000) if w["ref"] == "human" else "iid%d" % (w["ref_bases"] // 1_000_000)
    d = os.path.join(os.environ.get("YAHA_BENCH_CACHE", tempfile.gettempdir()), f"yaha_b200_bench_{tag}")
    os.makedirs(d, exist_ok=True)
    return d


_REF_CACHE = {}


def reference_bases(workload: str):
    """The workload's reference as (one uint8 character array, sequence bounds); deterministic (the last one made is kept)."""
    from yaha_b200 import synth
    w = WORKLOADS[workload]
    key = (w["ref"], w["ref_bases"])
    if key not in _REF_CACHE:
        _REF_CACHE.clear()
        if w["ref"] == "human":
            _REF_CACHE[key] = synth.human_like_reference(w["ref_bases"], 24, 2024)
        else:
            ref = synth.random_reference(w["ref_bases"], 12345)
            _REF_CACHE[key] = (ref, np.array([0, len(ref)], dtype=np.int64))
    return _REF_CACHE[key]


def make_reference(workload: str, with_fasta: bool = True):
    """Reference .nib2 (and FASTA, for `yaha -g`) on disk (cached); returns (dir, Nib2)."""
    from yaha_b200 import refio, synth
    d = cache_dir(workload)
    fa, nib = os.path.join(d, "ref.fa"), os.path.join(d, "ref.nib2")
    if not os.path.exists(nib):
        ref, bounds = reference_bases(workload)
        seqs = [(f"chr{k + 1}", ref[bounds[k]:bounds[k + 1]]) for k in range(len(bounds) - 1)]
        if with_fasta and len(ref) <= 500_000_000:                 # (human scale: the index is built on the device, no FASTA needed)
            synth.write_fasta(fa + ".tmp", seqs)
            os.replace(fa + ".tmp", fa)
        img = refio.build_nib2(seqs)
        with open(nib + ".tmp", "wb") as f:
            f.write(img)
        os.replace(nib + ".tmp", nib)
    return d, refio.load_nib2(nib)


def index_path_of(workload: str, d: str) -> str:
    from yaha_b200 import refio
    return os.path.join(d, refio.index_file_name("ref", 15, 1, WORKLOADS[workload]["index_h"]))


def ensure_device_built_index(workload: str, d: str, nib, device: int) -> str:
    """Index file in the reference's format, built ON THE DEVICE (ya_open_build: bit-identical to `yaha -g -L 15 -S 1 -H <h>`,
    tests/test_gpu_parity.py, tests/test_gpu_configs.py), written once per cache directory."""
    import yaha_b200
    from yaha_b200 import refio
    path = index_path_of(workload, d)
    if not os.path.exists(path):
        h = WORKLOADS[workload]["index_h"]
        al = yaha_b200.Aligner(nib, None, yaha_b200.Params.defaults(word_len=15, max_hits=min(650, h)), device=device, build_max_hits=h)
        refio.write_index(path + ".tmp", al.download_index(max_hits=h))
        os.replace(path + ".tmp", path)
        al.close()
    return path


def make_reads(workload: str, rank: int):
    """(name, char array) reads of this rank's shard; deterministic per rank."""
    from yaha_b200 import synth
    w = WORKLOADS[workload]
    ref, _ = reference_bases(workload)
    if w["reads"] == "sv":
        return list(synth.sv_reads(ref, w["n_reads"], w["read_len"], 4242 + 1000 * rank))
    return list(synth.simulate_reads(ref, w["n_reads"], w["read_len"], w["err"], 777 + 1000 * rank))


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region.  NVML in-process (one query costs well under a
    millisecond); falls back to spawning nvidia-smi, which is slow enough to disturb a 20 ms step, so less often."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.max_mhz = gpu, [], False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu))
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(gpu: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if gpu < len(ids) and ids[gpu].isdigit():
                return int(ids[gpu])
        return gpu

    def run(self):
        if self.nvml is not None:
            nv = self.nvml
            while not self.stop_flag:
                try:
                    mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((mhz, mask))
                except Exception:
                    pass
                time.sleep(0.05)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                if len(r) >= 6:
                    self.max_mhz = int(float(r[1]))
                    mask = sum(b for b, v in zip(bits, r[2:6]) if v.lower().startswith("active"))
                    self.rows.append((int(float(r[0])), mask))
            except Exception:
                pass
            time.sleep(1.0)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        seen = 0
        for _, m in self.rows:
            seen |= m
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": [n for b, n in self.REASONS.items() if seen & b],
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


_BARRIER_SEQ = [0]


def start_barrier_env(d: str, rank: int, world: int) -> dict:
    """N > 1: the ranks' host processes wait for each other once their index is resident (YA_START_BARRIER), so that every
    rank's timed passes run while all the others run theirs -- not while the others still upload 5 GB of index."""
    if world <= 1:
        return {}
    _BARRIER_SEQ[0] += 1
    bdir = os.path.join(d, f"barrier_{_BARRIER_SEQ[0]}")
    os.makedirs(bdir, exist_ok=True)
    try:
        os.remove(os.path.join(bdir, f"ready.{rank}"))
    except OSError:
        pass
    return {"YA_START_BARRIER": f"{bdir}:{rank}:{world}"}


def run_host(idx_path, reads_path, out_path, flags, threads, device, passes, batch, pipes, replay=False, tpp=0, env=None):
    """Run the product's host program (the call a user makes) and return its per-pass stats."""
    host = os.path.join(ROOT, "yaha_b200", "yaha_b200_host")
    if not os.path.exists(host):
        raise SystemExit("yaha_b200/yaha_b200_host is missing: run __graft_entry__.build() first (no CPU fallback)")
    cmd = [host, "-x", idx_path, "-q", reads_path, "-osh", out_path, "-t", str(threads), "-dev", str(device),
           "-passes", str(passes), "-batch", str(batch), "-pipes", str(pipes), "-tpp", str(tpp)] + (["-replay"] if replay else []) + flags
    p = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    if p.returncode != 0:
        raise SystemExit("yaha_b200_host failed:\n" + p.stderr[-3000:])
    return [json.loads(l) for l in p.stderr.splitlines() if l.startswith('{"pass"')]


def run_ours(args):
    import torch
    import torch.distributed as dist
    import yaha_b200
    from yaha_b200 import refio, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload
    W = WORKLOADS[wl]
    n_reads, rl, err = W["n_reads"], W["read_len"], W["err"]
    t0 = time.time()
    if rank == 0:
        d, nib = make_reference(wl)
        # index: built on the device (bit-identical to `yaha -g`), written in the reference's file format for both arms
        ensure_device_built_index(wl, d, nib, local)
    if world > 1:
        dist.barrier()
    d, nib = make_reference(wl)
    idx_path = index_path_of(wl, d)
    t_setup = time.time() - t0

    reads = make_reads(wl, rank)
    reads_path = os.path.join(d, f"reads_rank{rank}.fa")
    synth.write_reads(reads_path, reads)
    out_path = os.path.join(d, f"out_rank{rank}.sam")
    ncores = os.cpu_count() or 1
    threads = max(1, ncores // world)
    # a pipeline is a thread too (it drives the device and naps between polls): a rank that owns fewer cores than the default
    # 8 pipelines runs one pipeline per core (the configuration tools/sweep_fewcores.sh measured: 4 cores, 4 pipelines)
    pipes, e2e_pipes = (min(p, max(2, threads)) for p in (args.pipes, args.e2e_pipes))
    # ... and its device-driving threads nap 50 us between stream polls instead of 20 (YA_NAP_US): with 4 cores per GPU the job is
    # bound by host CPU, not by latency, and every poll is a wake-up taken from a worker
    few_cores_env = {"YA_NAP_US": os.environ.get("YA_NAP_US", "50")} if threads <= 4 else {}

    # roofline denominators measured live on this GPU: INT32 issue rate and the HBM random-gather rate
    # over the real 4 GiB starting-offset table (index rebuilt on the device for this, ~1 s)
    probe_al = yaha_b200.Aligner(nib, None, yaha_b200.Params.defaults(word_len=15), device=local, build_max_hits=W["index_h"])
    int_add, int_mix = probe_al.int32_peak()
    gather_peak = probe_al.gather_peak()
    probe_al.close()

    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # run A: the job exactly as a user runs it (FASTA parsed, SAM written) -> e2e.  Tuned for the latency
    # of a 20 K-read job: small batches on 8 overlapping pipelines served by one pool of worker threads.
    e2e_tpp = 0                      # shared worker pool of `threads` workers serves every pipeline
    extra_warm = 5                   # (page-locked / device scratch of 8 pipelines reaches its final size in the first passes)
    def benv(extra=None):
        # (every rank clears its own ready file, then all meet, then the host processes meet again with their index resident)
        e = dict(few_cores_env, **(extra or {}), **start_barrier_env(d, rank, world))
        if world > 1:
            dist.barrier()
        return e
    stats_a = run_host(idx_path, reads_path, out_path, REF_FLAGS[wl], threads, local, extra_warm + args.warmup + args.steps,
                       args.e2e_batch, e2e_pipes, tpp=e2e_tpp, env=benv())
    timed_a = stats_a[extra_warm + args.warmup:]
    if os.environ.get("YAHA_BENCH_DUMP_STATS"):              # per-rank, per-pass stats of the e2e run (debugging stragglers)
        with open(os.path.join(os.environ["YAHA_BENCH_DUMP_STATS"], f"stats_a_rank{rank}.json"), "w") as f:
            json.dump(stats_a, f)
    assert len(timed_a) == args.steps, (len(stats_a), args.warmup, args.steps)
    el_e2e = sum(s["align_s"] for s in timed_a)
    # run B: parsed reads replayed from host memory, SAM formatted but not written -> value, stage times and the kernel
    # rooflines.  Two pipelines x 10 000-read batches: every extension launch carries ~20 K jobs (0.7 of the INT32 roofline per
    # launch).  The latency-tuned configuration of the e2e run (four pipelines x 5000 reads) is replayed as run D and reported
    # beside it (`value_at_e2e_config`: more reads/s, launches half the size); run C times the kernels alone on whole-shard batches.
    stats_b = run_host(idx_path, reads_path, out_path + ".replay", REF_FLAGS[wl], threads, local, 1 + extra_warm + args.warmup + args.steps,
                       args.batch, pipes, replay=True, env=benv())
    timed = stats_b[1 + extra_warm + args.warmup:]
    assert len(timed) == args.steps
    el_res = sum(s["align_s"] for s in timed)
    stats_d = run_host(idx_path, reads_path, out_path + ".replay", REF_FLAGS[wl], threads, local, 1 + extra_warm + args.warmup + args.steps,
                       args.e2e_batch, e2e_pipes, replay=True, env=benv())[1 + extra_warm + args.warmup:]
    d_el = sum(s["align_s"] for s in stats_d)
    d_cells, d_ms, d_union = (sum(s[k] for s in stats_d) for k in ("ext_cells", "dev_ms_ext", "dev_ms_ext_union"))
    # run C (not part of value/e2e): the same job with the whole shard as ONE batch on one pipeline, DP rounds in
    # lock step -> every extension launch carries the workload's ~40 K jobs and no other pipeline shares the SMs.
    # This is the timed region the kernel rooflines are quoted on; run B's small launches are reported beside it.
    stats_c = run_host(idx_path, reads_path, out_path + ".replay", REF_FLAGS[wl], threads, local, 1 + args.warmup + args.steps,
                       n_reads, 1, replay=True, env=benv({"YA_COALESCE_US": "20000"}))[1 + args.warmup:]
    iso_cells, iso_ms, iso_n = (sum(s[k] for s in stats_c) for k in ("ext_cells", "dev_ms_ext", "ext_launches"))
    iso_gcups = iso_cells / (iso_ms * 1e-3) / 1e9 if iso_ms > 0 else 0.0
    iso_lookup_ms, iso_probes = sum(s["dev_ms_lookup"] for s in stats_c), sum(s["probes"] for s in stats_c)
    iso_pps = iso_probes / (iso_lookup_ms * 1e-3) if iso_lookup_ms > 0 else 0.0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    per_rank = None
    if world > 1:
        mine_t = torch.tensor([el_e2e, el_res], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(allt, mine_t)
        per_rank = {"e2e_ms_per_step": [round(float(t[0]) / args.steps * 1e3, 3) for t in allt],
                    "value_ms_per_step": [round(float(t[1]) / args.steps * 1e3, 3) for t in allt]}
        el_e2e, el_res = max(float(t[0]) for t in allt), max(float(t[1]) for t in allt)

    def tot(k):
        return sum(s[k] for s in timed)

    total_reads = n_reads * world * args.steps
    value = total_reads / el_res
    e2e = total_reads / el_e2e
    cells, ms_dp, ms_seed, ms_tb = tot("dp_cells"), tot("dev_ms_dp"), tot("dev_ms_seed"), tot("dev_ms_traceback")
    gcups = cells / (ms_dp * 1e-3) / 1e9 if ms_dp > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650"
    ext_cells, ms_ext, ext_launches = tot("ext_cells"), tot("dev_ms_ext"), tot("ext_launches")
    ext_gcups = ext_cells / (ms_ext * 1e-3) / 1e9 if ms_ext > 0 else 0.0
    achieved_giops = ext_gcups * INT_OPS_PER_CELL_EXT
    seed_bytes = 8.0 * tot("probes") + 20.0 * tot("hits") + 12.0 * tot("frags_all")
    seed_gbs = seed_bytes / (ms_seed * 1e-3) / 1e9 if ms_seed > 0 else 0.0
    ms_lookup = tot("dev_ms_lookup")
    probes_per_s = tot("probes") / (ms_lookup * 1e-3) if ms_lookup > 0 else 0.0
    k2_bytes = 20.0 * tot("hits") + 12.0 * tot("frags_all")
    k2_gbs = k2_bytes / ((ms_seed - ms_lookup) * 1e-3) / 1e9 if ms_seed > ms_lookup else 0.0
    iso_k2_ms = sum(s["dev_ms_seed"] - s["dev_ms_lookup"] for s in stats_c)
    iso_k2_gbs = sum(20.0 * s["hits"] + 12.0 * s["frags_all"] for s in stats_c) / (iso_k2_ms * 1e-3) / 1e9 if iso_k2_ms > 0 else 0.0
    # bytes that cross PCIe per step in the e2e run: the reads as characters + ids + offsets up, SAM text + offsets + status down
    # (plus the call-by-call traffic of the few reads handed back: their codes up, fragments / DP answers down -- counted at 2 KB each)
    handed = timed_a[-1].get("reads_handed_back", 0)
    h2d = int(sum(len(s) for _, s in reads) + sum(len(n) for n, _ in reads) + 12 * (len(reads) + 1) + 1024 * handed)
    d2h = int(timed_a[-1].get("device_text_bytes", 0) + 9 * (len(reads) + 1) + 2048 * handed)

    # DRAM bytes of the two dominant kernels from the committed `ncu --set full` captures, scaled from the
    # captured launch to this run's average launch (per cell / per probe)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    t_ext, t_seed = traffic.get("dp_ext_packed_kernel"), traffic.get("seed_count_kernel")

    line = {
        "metric": METRIC,
        "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": el_res / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[wl], "reads_per_gpu": n_reads, "read_len": rl, "error": err,
                   "flags": REF_FLAGS[wl], "host_threads_per_gpu": threads,
                   "l2": "inputs larger than L2 (4.3 GB index gathers; reads re-uploaded every step)",
                   "value_run": {"batch_reads": args.batch, "pipelines_per_gpu": pipes, "worker_pool_threads": threads},
                   "e2e_run": {"batch_reads": args.e2e_batch, "pipelines_per_gpu": e2e_pipes, "worker_pool_threads": threads},
                   "stream_poll_nap_us": int(few_cores_env.get("YA_NAP_US", os.environ.get("YA_NAP_US", "20"))),
                   "value_excludes": "FASTA parsing and SAM fwrite (reads replayed from host memory; the 10 MB/step H2D of "
                                     "read codes is still inside); e2e includes everything",
                   "setup_s": round(t_setup, 2)},
        "gcups": ext_gcups, "gcups_kernel": "dp_ext_packed_kernel (banded X-drop extension, 92 % of all DP cells), value run",
        "gcups_isolated_run_c": iso_gcups,
        "gcups_all_dp_kernels_in_value_run": gcups, "dp_cells_per_step": cells // max(args.steps, 1), "dp_jobs_per_step": tot("dp_jobs") // args.steps,
        "dp_calls_per_step": tot("dp_rounds") // args.steps,
        "stage_ms_per_step": {"note": f"value run; device_* are CUDA-event spans on each pipeline's stream ({pipes} pipelines overlap on the device, so "
                                      "the spans overlap and include queueing); wall_host_logic is worker-pool busy time per thread",
                              "device_seed": ms_seed / args.steps, "device_dp_fill": ms_dp / args.steps,
                              "device_traceback": ms_tb / args.steps, "device_assemble_finish_format": tot("dev_ms_finish") / args.steps,
                              "wall_seed_call": tot("seed_wall_s") / args.steps * 1e3, "wall_dp_calls": tot("dp_wall_s") / args.steps * 1e3,
                              "wall_host_logic": tot("host_wall_s") / args.steps * 1e3, "wall_parse": tot("read_parse_s") / args.steps * 1e3,
                              "wall_upload": tot("upload_s") / args.steps * 1e3, "wall_write": tot("write_s") / args.steps * 1e3},
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_timed_step": [round(s["align_s"] * 1e3, 2) for s in timed_a],
                "ms_per_untimed_step": [round(s["align_s"] * 1e3, 2) for s in stats_a[:extra_warm + args.warmup]]},
        "value_ms_per_timed_step": [round(s["align_s"] * 1e3, 2) for s in timed],
        "value_at_e2e_config": {"what": f"the same replay with the e2e run's configuration ({args.e2e_batch}-read batches on {e2e_pipes} pipelines): "
                                        "higher throughput, extension launches half the size",
                                "value_this_rank": n_reads * args.steps / d_el, "unit": "reads/s", "ms_per_step": d_el / args.steps * 1e3,
                                "roofline_frac_per_launch": d_cells / (d_ms * 1e-3) / 1e9 * INT_OPS_PER_CELL_EXT / int_add if d_ms > 0 and int_add else None,
                                "roofline_frac_device_level": d_cells / (d_union * 1e-3) / 1e9 * INT_OPS_PER_CELL_EXT / int_add if d_union > 0 and int_add else None},
        "gpu_launches": int(tot("launches")), "gpu_launches_per_step": int(tot("launches")) // args.steps,
        "reads_finished_on_device_per_step": timed[-1].get("reads_finished_on_device"), "reads_handed_back_per_step": timed[-1].get("reads_handed_back"),
        # frac is the figure of the run that produces `value` (the product's own launches); the same kernel timed alone on
        # whole-shard launches (run C) is reported beside it as `isolated`
        "roofline": {"bound": "int32-issue", "kernel": "dp_ext_packed_kernel",
                     "timed_region": f"value run: bulk extension launches (>= 4096 jobs) of the {pipes} pipelines, CUDA events on the launching stream",
                     "launches_timed": int(ext_launches), "avg_launch_ms": ms_ext / max(ext_launches, 1),
                     "cells_per_launch": ext_cells / max(ext_launches, 1),
                     "achieved": achieved_giops, "peak": int_add, "unit": "GIOP/s",
                     "frac": achieved_giops / int_add if int_add else None,
                     "traffic": t_ext["dram_bytes"] / t_ext["cells"] * (ext_cells / max(ext_launches, 1)) if t_ext else None,
                     "traffic_unit": "bytes per launch (dram__bytes_read+write of the ncu capture in profiles/ncu_traffic.json, per cell x cells_per_launch)",
                     "algorithmic_bytes_per_launch": 1.0 * ext_cells / max(ext_launches, 1),
                     "peak_source": "ya_measure_int32_peak: dependent-free IADD/LOP3 stream on all SMs, measured live (not in MEASURED_PEAKS.json, "
                                    "which holds HBM and bf16 figures only)",
                     "peak_dp_mix": int_mix, "ops_per_cell": INT_OPS_PER_CELL_EXT, "gcups": ext_gcups,
                     "gcups_roof": int_add / INT_OPS_PER_CELL_EXT,
                     "device_level": {"what": "the same cells over the UNION of the launches' spans on the device's time axis: the pipelines "
                                              "overlap their launches, which stretches every one of them (per-launch frac above) while the "
                                              "device as a whole does more",
                                      "union_ms_per_step": tot("dev_ms_ext_union") / args.steps,
                                      "gcups": ext_cells / (tot("dev_ms_ext_union") * 1e-3) / 1e9 if tot("dev_ms_ext_union") > 0 else None,
                                      "frac": ext_cells / (tot("dev_ms_ext_union") * 1e-3) / 1e9 * INT_OPS_PER_CELL_EXT / int_add
                                              if tot("dev_ms_ext_union") > 0 and int_add else None},
                     "hardware_utilisation_ncu": {"alu_pipe_pct": (t_ext or {}).get("alu_pipe_pct"), "issue_active_pct": (t_ext or {}).get("issue_active_pct"),
                                                  "thread_instructions_per_cell": (t_ext or {}).get("thread_inst_per_cell"),
                                                  "note": "frac divides ALGORITHMIC ops (33/cell, SURVEY 8d) by an instruction rate; DPX fuses 2-3 of them "
                                                          "per instruction, so the hardware figure is the ncu pipe utilisation of the committed capture"},
                     "isolated": {"timed_region": "run C: the same job as one batch on one pipeline, DP rounds in lock step (every bulk extension "
                                                  "launch carries the workload's whole first round)",
                                  "launches_timed": int(iso_n), "avg_launch_ms": iso_ms / max(iso_n, 1),
                                  "cells_per_launch": iso_cells / max(iso_n, 1), "gcups": iso_gcups,
                                  "achieved": iso_gcups * INT_OPS_PER_CELL_EXT,
                                  "frac": iso_gcups * INT_OPS_PER_CELL_EXT / int_add if int_add else None}},
        "roofline_k2": {"bound": "hbm", "kernel": "hits -> fragments -> regions (expand_hits, segmented sort, fragment / region / survivor scans)",
                        "timed_region": "value run: CUDA-event span of ya_seed_frags minus the seed_count_kernel span",
                        "algorithmic_bytes": "4 B/hit ROA read + 16 B/hit (key written and read once) + 12 B/fragment (SURVEY 8d)",
                        "achieved": k2_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k2_gbs / hbm_peak if hbm_peak else None,
                        "hits_per_step": tot("hits") / args.steps, "frags_per_step": tot("frags_all") / args.steps,
                        "ms_per_step": (ms_seed - ms_lookup) / args.steps, "peak_source": hbm_src,
                        "isolated": {"timed_region": "run C", "achieved": iso_k2_gbs, "frac": iso_k2_gbs / hbm_peak if hbm_peak else None,
                                     "ms_per_step": iso_k2_ms / max(len(stats_c), 1)}},
        "roofline_seed": {"bound": "hbm", "kernel": "seed_count_kernel (k-mer -> starting-offset gather, Query.c:391)",
                          "timed_region": "run C (see roofline.timed_region)",
                          "achieved": 8.0 * iso_pps / 1e9, "peak": hbm_peak, "unit": "GB/s",
                          "frac": 8.0 * iso_pps / 1e9 / hbm_peak, "traffic": t_seed["dram_bytes"] / t_seed["probes"] * (iso_probes / max(len(stats_c), 1)) if t_seed else None,
                          "peak_source": hbm_src,
                          "algorithmic_bytes_per_probe": 8, "probes_per_s": iso_pps,
                          "random_gather_peak_per_s": gather_peak,
                          "frac_of_random_gather_peak": iso_pps / gather_peak if gather_peak else None,
                          "in_value_run_probes_per_s": probes_per_s,
                          "note": "a probe is one independent 32 B-sector DRAM miss in a 4 GiB table; the binding limit is the HBM "
                                  "random-access rate (measured live by ya_measure_gather_peak), not streaming bandwidth",
                          "whole_stage_GBps_in_value_run": seed_gbs, "whole_stage": "seed_count + expand + segmented sort + fragment/region scans, "
                                                                        "8 B/probe + 20 B/hit + 12 B/fragment"},
        "clocks": sampler.summary(),
    }
    if per_rank:
        line["per_rank"] = per_rank
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()                  # (the other ranks are done: their GPUs are free for the in-process run below)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(wl, d, idx_path, reads, min(n_reads, args.cpu_sample), out_path, min(n_reads, args.t1_sample))
        if world > 1 and not args.no_inprocess:
            line["inprocess"] = inprocess_multi_gpu(args, wl, d, idx_path, world, ncores)
        print(json.dumps(line), flush=True)


def inprocess_multi_gpu(args, wl, d, idx_path, world, ncores):
    """The product's own multi-GPU path: ONE `yaha_b200_host -gpus N` process on one file holding the reads of all
    N ranks -- index uploaded once and replicated GPU-to-GPU (ya_open_peer), batches of the one query stream dealt to the
    devices, SAM written in input order.  Its SAM must equal the N per-rank outputs one after the other."""
    from yaha_b200 import synth
    time.sleep(1.0)                                   # let the other ranks' processes leave their GPUs
    q = os.path.join(d, f"reads_all{world}.fa")
    with open(q, "wb") as f:
        for r in range(world):
            f.write(open(os.path.join(d, f"reads_rank{r}.fa"), "rb").read())
    out = os.path.join(d, f"out_all{world}.sam")
    host = os.path.join(ROOT, "yaha_b200", "yaha_b200_host")
    n_pass = 3 + args.warmup + args.steps
    cmd = [host, "-x", idx_path, "-q", q, "-osh", out, "-t", str(ncores), "-gpus", str(world), "-passes", str(n_pass),
           "-batch", str(args.e2e_batch), "-pipes", str(args.e2e_pipes)] + REF_FLAGS[wl]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        return {"error": p.stderr[-1500:]}
    st = [json.loads(l) for l in p.stderr.splitlines() if l.startswith('{"pass"')]
    timed = st[-args.steps:]
    el = sum(s["align_s"] for s in timed)
    body = [l for l in open(out, "rb") if not l.startswith(b"@")]
    want = []
    for r in range(world):
        want += [l for l in open(os.path.join(d, f"out_rank{r}.sam"), "rb") if not l.startswith(b"@")]
    return {"what": f"one process, -gpus {world}, one query file with the {world} x {WORKLOADS[wl]['n_reads']} reads of all ranks, SAM written in input order",
            "value": sum(s["reads"] for s in timed) / el, "unit": "reads/s", "ms_per_step": el / len(timed) * 1e3,
            "ms_per_timed_step": [round(s["align_s"] * 1e3, 2) for s in timed],
            "sam_equals_the_per_rank_outputs_in_order": body == want, "sam_records": len(body),
            "index_upload_s": st[-1].get("index_upload_s"), "index_peer_copies_s": st[-1].get("index_peer_copies_s"),
            "peer_copies_direct": st[-1].get("peer_copies_direct"), "open_s": st[-1].get("open_s"),
            "reads_finished_on_device": timed[-1].get("reads_finished_on_device"), "reads_handed_back": timed[-1].get("reads_handed_back")}


def ensure_index_file(wl: str, d: str) -> str:
    """Index file on disk for the reference binary: the one our arm built on the device if it is there, else built by the
    reference itself (`yaha -g`; only for references small enough to have a FASTA in the cache directory)."""
    path = index_path_of(wl, d)
    if not os.path.exists(path):
        yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
        subprocess.check_call([yaha, "-g", "ref.fa", "-L", "15", "-S", "1", "-H", str(WORKLOADS[wl]["index_h"])], cwd=d,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return path


def time_reference(wl: str, d: str, index_path: str, reads_path: str, threads: int, out: str):
    yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
    cmd = [yaha, "-x", index_path, "-q", reads_path, "-osh", out, "-t", str(threads)] + REF_FLAGS[wl]
    t = time.perf_counter()
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t


def reference_rate(wl, d, idx, reads, ncores, tag):
    """reads/s of the unmodified reference: whole-program wall time, and the align phase alone
    (wall minus the wall of the same command on a 1-read file = index/genome mmap + page touch)."""
    from yaha_b200 import synth
    qf, q1 = os.path.join(d, f"{tag}_{len(reads)}.fa"), os.path.join(d, f"{tag}_one.fa")
    synth.write_reads(qf, reads)
    synth.write_reads(q1, reads[:1])
    out = os.path.join(d, f"{tag}_ref_out.sam")
    time_reference(wl, d, idx, q1, ncores, out)                 # warm the page cache
    t_load = min(time_reference(wl, d, idx, q1, ncores, out) for _ in range(2))
    t_full = time_reference(wl, d, idx, qf, ncores, out)
    return t_full, t_load, out


def cpu_baseline(wl, d, idx, reads, sample, my_sam, t1_sample=0):
    ncores = os.cpu_count() or 1
    yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
    if not os.path.exists(yaha):
        return {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "oracle/_ref/yaha missing"}
    sub = reads[:sample]
    t_full, t_load, ref_sam = reference_rate(wl, d, idx, sub, ncores, "cpu")
    res = {"value": len(sub) / max(t_full - t_load, 1e-9), "unit": "reads/s", "cores": ncores, "kind": "reference",
           "whole_program_reads_per_s": len(sub) / t_full, "index_load_s": round(t_load, 3),
           "sample": f"{len(sub)} reads of the workload, unmodified yaha -t {ncores}; value = reads / (wall - wall of a 1-read run); "
                     f"whole-program wall {t_full:.2f} s"}
    # SAM identity on every run: the reference's records for the sampled reads against the product's records for the
    # same reads (multiset: `yaha -t N` emits reads in completion order), and once in order against `yaha -t 1`
    names = {n for n, _ in sub}
    mine = [l for l in open(my_sam) if not l.startswith("@PG")]           # (@PG echoes file names and -t, AlignOutput.c:50-60)
    if len(sub) != len(reads):
        mine = [l for l in mine if l.startswith("@") or l.split("\t", 1)[0] in names]
    a = sorted(l for l in open(ref_sam) if not l.startswith("@PG"))
    res["sam_identical_to_reference"] = (a == sorted(mine))
    res["sam_records"] = len(a)
    if t1_sample > 0:
        q1 = os.path.join(d, "cpu_t1.fa")
        from yaha_b200 import synth
        synth.write_reads(q1, reads[:t1_sample])
        out1 = os.path.join(d, "cpu_t1.sam")
        time_reference(wl, d, idx, q1, 1, out1)
        want = [l for l in open(out1) if not l.startswith("@PG")]
        n1 = {n for n, _ in reads[:t1_sample]}
        got = [l for l in mine if l.startswith("@") or l.split("\t", 1)[0] in n1] if t1_sample != len(reads) else mine
        res["sam_identical_in_order_to_t1"] = (want == got)
        res["sam_records_t1"] = len(want)
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    n_reads = WORKLOADS[wl]["n_reads"]
    yaha = os.path.join(ROOT, "oracle", "_ref", "yaha")
    if not os.path.exists(yaha):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/yaha not built (reference sources absent)"}))
        return
    d, nib = make_reference(wl)
    idx = ensure_index_file(wl, d)
    reads = make_reads(wl, 0)
    sample = reads[:min(n_reads, args.cpu_sample)]
    ncores = os.cpu_count() or 1
    times, load = [], None
    for k in range(args.warmup + args.steps):
        t_full, t_load, _ = reference_rate(wl, d, idx, sample, ncores, "refarm")
        if k >= args.warmup:
            times.append(max(t_full - t_load, 1e-9))
            load = t_load
    tot = sum(times)
    v = len(sample) * len(times) / tot
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[wl], "flags": REF_FLAGS[wl], "threads": ncores,
                       "timing": "wall of `yaha -t <cores>` minus wall of the same command on a 1-read file (index mmap + page touch)",
                       "index_load_s": load},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": ncores, "kind": "reference",
                             "sample": f"{len(sample)} reads per step, unmodified yaha -t {ncores}"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=1 << 30, help="reads of the workload the reference arm / cpu_baseline aligns (default: all)")
    ap.add_argument("--t1-sample", type=int, default=20000, help="reads aligned once with `yaha -t 1` for the in-order SAM comparison")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inprocess", action="store_true", help="N > 1: skip the one-process -gpus N measurement on rank 0")
    ap.add_argument("--batch", type=int, default=10000, help="reads per device batch in the value run")
    ap.add_argument("--pipes", type=int, default=2, help="concurrent batch pipelines per GPU in the value run")
    ap.add_argument("--e2e-batch", type=int, default=5000, help="reads per device batch in the e2e run")
    ap.add_argument("--e2e-pipes", type=int, default=4, help="pipelines per GPU in the e2e run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
