"""Every BASELINE.json config, end to end on the GPU: the product's host program against the UNMODIFIED reference
(`oracle/_ref/yaha -t 1`), SAM compared line by line IN ORDER (everything except @PG).  Shapes and flags are bench.py's
WORKLOADS; read counts are reduced where the reference's single thread would take minutes.

cfg4 / cfg5 use the human-scale reference (3.1 Gbp in 24 sequences, Alu-like family, N runs) with an index built ON THE
DEVICE with -H 650, i.e. through the down-sampling of over-full k-mers (Index.c:271-315).  The size drops to 0.4 Gbp only if
the box lacks the RAM / disk for the 16.7 GB index file (stated in the test output)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

HOST = os.path.join(ROOT, "yaha_b200", "yaha_b200_host")
REF = os.path.join(ROOT, "oracle", "_ref", "yaha")


def _pick_human_size(tmp):
    try:
        ram = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        ram = 0
    disk = shutil.disk_usage(tmp).free
    return "3.1" if (ram >= 60 << 30 and disk >= 40 << 30) else "0.4"


@pytest.fixture(scope="module")
def bench(tmp_path_factory):
    cache = os.environ.get("YAHA_BENCH_CACHE") or str(tmp_path_factory.mktemp("cfg"))
    os.environ["YAHA_BENCH_CACHE"] = cache
    os.environ.setdefault("YAHA_BENCH_GBP", _pick_human_size(cache))
    sys.modules.pop("bench", None)
    import bench as B
    yield B
    if "YAHA_KEEP_CACHE" not in os.environ:
        shutil.rmtree(cache, ignore_errors=True)


def _sam(path):
    return [l for l in open(path) if not l.startswith("@PG")]


# reads compared per config (all of them where `yaha -t 1` needs seconds)
N_CHECK = {"cfg1": 10_000, "cfg2": 100_000, "cfg3": 20_000, "cfg4": 250, "cfg5": 4_000}


@pytest.mark.parametrize("wl", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_baseline_config_sam_identical_in_order(bench, wl):
    from yaha_b200 import synth
    B = bench
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/yaha not built")
    W = dict(B.WORKLOADS[wl])
    d, nib = B.make_reference(wl, with_fasta=False)
    idx = B.ensure_device_built_index(wl, d, nib, 0)
    n = min(N_CHECK[wl], W["n_reads"])
    B.WORKLOADS[wl]["n_reads"] = n
    try:
        reads = B.make_reads(wl, 0)
    finally:
        B.WORKLOADS[wl]["n_reads"] = W["n_reads"]
    q = os.path.join(d, f"{wl}_check.fa")
    synth.write_reads(q, reads)
    mine, want = os.path.join(d, f"{wl}_mine.sam"), os.path.join(d, f"{wl}_ref.sam")
    p = subprocess.run([HOST, "-x", idx, "-q", q, "-osh", mine, "-t", str(os.cpu_count() or 4)] + W["flags"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    p = subprocess.run([REF, "-x", idx, "-q", q, "-osh", want, "-t", "1"] + W["flags"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    a, b = _sam(mine), _sam(want)
    print(f"{wl}: reference {W['ref_bases'] / 1e9:.2f} Gbp, {n} reads, {sum(1 for l in b if not l.startswith('@'))} records")
    assert len(b) > n // 4
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, (wl, i, x[:200], y[:200])
    assert len(a) == len(b)


def test_down_sampled_index_equals_the_reference_build(bench, tmp_path):
    """ya_open_build with -H 650 on a repeat-rich reference (40 Mbp, one Alu-like copy per kbp: hundreds of k-mers far
    above the cap) against the file `yaha -g -H 650` writes: the sampled lists must be the reference's, entry for entry."""
    import hashlib
    import yaha_b200
    from yaha_b200 import refio, synth
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/yaha not built")
    ref, bounds = synth.human_like_reference(40_000_000, 5, 7, alu_sites=40000)
    seqs = [(f"chr{k + 1}", ref[bounds[k]:bounds[k + 1]]) for k in range(len(bounds) - 1)]
    synth.write_fasta(str(tmp_path / "ref.fa"), seqs)
    subprocess.check_call([REF, "-g", "ref.fa", "-L", "15", "-S", "1", "-H", "650"], cwd=str(tmp_path), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    theirs = tmp_path / refio.index_file_name("ref", 15, 1, 650)
    nib = refio.load_nib2(str(tmp_path / "ref.nib2"))                 # written by the reference
    al = yaha_b200.Aligner(nib, None, yaha_b200.Params.defaults(word_len=15), device=0, build_max_hits=650)
    idx = al.download_index(max_hits=650)
    al.close()
    assert int(np.max(np.diff(idx.so.astype(np.int64)))) == 650
    refio.write_index(str(tmp_path / "mine.idx"), idx)

    def sha(p):
        h = hashlib.sha256()
        with open(p, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        return h.hexdigest()
    assert sha(tmp_path / "mine.idx") == sha(theirs)
