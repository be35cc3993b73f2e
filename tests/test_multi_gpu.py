"""The product's own multi-GPU path on hardware: one host process, `-gpus 2`, index uploaded once and
replicated to the second device by ya_open_peer (peer access enabled: a direct copy over NVLink), batches of ONE query stream
dealt to both devices, SAM in input order -- identical to the reference's.  Skipped on a box with one GPU."""
import json
import os
import subprocess

import pytest

import hostcases as H
import support as S

pytestmark = pytest.mark.gpu
HOST = os.path.join(S.ROOT, "yaha_b200", "yaha_b200_host")


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("env", [{}, {"YA_FUSED": "0"}], ids=["align_batch", "fibers"])
def test_two_gpus_one_process_same_sam(small, tmp_path, env):
    want = H.expected(small, "out_bw10.sam.gz")
    out = str(tmp_path / "o.sam")
    # small batches so that both devices get work; strands above 512 hits take the big segmented sort on BOTH devices
    cmd = H.command(HOST, small, "reads.fa", "-osh", out, ["-BW", "10", "-G", "100", "-gpus", "2", "-batch", "40", "-pipes", "2"], threads=4)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, YAHA_B200_STATS="1", **env))
    assert p.returncode == 0, p.stderr[-2000:]
    assert H.sam_lines(open(out).read()) == want
    st = [json.loads(l) for l in p.stderr.splitlines() if l.startswith('{"pass"')][-1]
    assert st["gpus"] == 2 and st["peer_copies_direct"] == 1, st


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_gpus_repeat_rich_strands(small, tmp_path):
    """Reads with more than 512 seed hits per strand (the 64 KB shared-memory sort, whose opt-in is per device) on both
    devices: chimera.fa has them; the SAM must be the reference's."""
    want = H.expected(small, "out_chimera.sam.gz")
    out = str(tmp_path / "c.sam")
    cmd = H.command(HOST, small, "chimera.fa", "-osh", out, ["-gpus", "2", "-batch", "7", "-pipes", "2"], threads=4)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert H.sam_lines(open(out).read()) == want
