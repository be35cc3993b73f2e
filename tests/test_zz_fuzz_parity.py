"""Differential fuzzing against the unmodified reference as part of the suite (tools/fuzz_parity.py): random references with
repeat families / tandem repeats / N runs, random index geometry (-L/-S/-H), read mixes, output formats and flag sets; the SAM of
`oracle/_ref/yaha -t 1` and of the host program must agree line by line.  A dozen seeds here (CPU: the host program on the mock
of the ABI, i.e. the per-read sources the kernels compile; GPU: the CUDA library); hundreds more were run by hand
(profiles/r02_fuzz_gpu.log)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUZZ = os.path.join(ROOT, "tools", "fuzz_parity.py")
REF = os.path.join(ROOT, "oracle", "_ref", "yaha")
HOST = os.path.join(ROOT, "yaha_b200", "yaha_b200_host")

needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/yaha not built")


def _fuzz(args, tmp_path, min_compared):
    p = subprocess.run([sys.executable, FUZZ, "--keep", str(tmp_path / "fail")] + args, capture_output=True, text=True, timeout=1700)
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-2000:]
    m = re.match(r"(\d+) cases, (\d+) compared, (\d+) failed", tail)
    assert m, (tail, p.stderr[-2000:])
    assert int(m.group(3)) == 0 and p.returncode == 0, p.stdout[-4000:]
    assert int(m.group(2)) >= min_compared, p.stdout[-4000:]


@needs_ref
def test_fuzz_host_and_shared_sources_on_cpu(tmp_path):
    _fuzz(["--seeds", "100:110", "--jobs", "4", "--wordlens", "11,12"], tmp_path, 8)


@pytest.mark.gpu
@needs_ref
def test_fuzz_cuda_library(tmp_path):
    # (the seeds and -L list of the round's GPU fuzz run, profiles/r02_fuzz_gpu.log)
    _fuzz(["--seeds", "20000:20016", "--jobs", "8", "--binary", HOST, "--wordlens", "11,11,12,13"], tmp_path, 12)
