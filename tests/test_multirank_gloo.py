"""N>1 path on the CPU (world_size 2, gloo): reads shard by contiguous record ranges, every rank
aligns its shard independently (replicated index, no data-path collective), rank-ordered
concatenation of the shard outputs equals the single-process output, and the step time is the MAX
over ranks -- the same plumbing bench.py uses under torchrun.  Ranks run the oracle-backed mock host
(no GPU here)."""
import os
import subprocess
import sys

import pytest

import hostcases as H
import support as S

WORKER = r'''
import os, sys, subprocess, time
import torch, torch.distributed as dist
rank, world, workdir, mock, idx = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[6]
dist.init_process_group("gloo", rank=rank, world_size=world)
recs = open(os.path.join(workdir, "reads.fa")).read().split(">")[1:]
lo, hi = rank * len(recs) // world, (rank + 1) * len(recs) // world          # contiguous slice of the query file
shard = os.path.join(workdir, f"shard{rank}.fa")
open(shard, "w").write("".join(">" + r for r in recs[lo:hi]))
dist.barrier()
t = time.perf_counter()
subprocess.run([mock, "-x", idx, "-q", shard, "-osh", os.path.join(workdir, f"out{rank}.sam"), "-t", "1"], check=True, capture_output=True)
el = torch.tensor([time.perf_counter() - t], dtype=torch.float64)
dist.all_reduce(el, op=dist.ReduceOp.MAX)                                    # time = max over ranks
n = torch.tensor([hi - lo], dtype=torch.int64)
dist.all_reduce(n, op=dist.ReduceOp.SUM)                                     # value = all reads / that time
if rank == 0:
    open(os.path.join(workdir, "summary.txt"), "w").write(f"{int(n.item())} {float(el.item())}\n")
dist.destroy_process_group()
'''


def test_two_rank_sharding_preserves_output(small, tmp_path):
    mock = os.path.join(S.ROOT, "tests", "_build", "yaha_host_mock")
    subprocess.check_call(["make", "-s", "-C", os.path.join(S.ROOT, "tests", "mock"), "SAN="])
    work = str(tmp_path)
    open(os.path.join(work, "reads.fa"), "w").write(open(os.path.join(small.dir, "reads.fa")).read())
    script = os.path.join(work, "worker.py")
    open(script, "w").write(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", work, mock, small.idx_path, port]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    body = []
    for r in range(2):
        body += [l for l in H.sam_lines(open(os.path.join(work, f"out{r}.sam")).read()) if not l.startswith("@")]
    want = [l for l in H.expected(small, "out_bw5.sam.gz") if not l.startswith("@")]
    assert body == want
    n, el = open(os.path.join(work, "summary.txt")).read().split()
    assert int(n) == 635 and float(el) > 0
