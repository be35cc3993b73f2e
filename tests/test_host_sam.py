"""End-to-end drop-in check on the GPU: the batched host program (yaha_b200/host) with the CUDA
hot path behind the C ABI must write the same SAM as the unmodified reference `yaha -t 1` --
every line except @PG, in the same order -- for every golden flag set."""
import os
import subprocess

import pytest

import hostcases as H
import support as S

pytestmark = pytest.mark.gpu
HOST = os.path.join(S.ROOT, "yaha_b200", "yaha_b200_host")


@pytest.mark.parametrize("golden,reads,outflag,extra", H.CASES)
def test_sam_identical_to_reference(small, tmp_path, golden, reads, outflag, extra):
    out = str(tmp_path / "o.sam")
    p = subprocess.run(H.command(HOST, small, reads, outflag, out, extra), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    got, want = H.sam_lines(open(out).read()), H.expected(small, golden)
    diff = [(i, x, y) for i, (x, y) in enumerate(zip(got, want)) if x != y]
    assert not diff, diff[:2]
    assert len(got) == len(want)


@pytest.mark.parametrize("flags", H.FLAG_SWEEP, ids=lambda f: "".join(f))
def test_every_alignment_flag_gives_reference_sam(small, tmp_path, flags):
    """Every alignment flag of the reference's CLI, alone and combined, degenerate values included (band 0, X-drop 0, one hit
    per k-mer, band wider than the gap cap -- parameter sets the packed kernel does not serve): digest of the SAM equals the
    digest of the reference's (golden/small/flag_sweep.json)."""
    H.check_flag_sweep(HOST, small, str(tmp_path / "o.sam"), flags, threads=3)


def test_sam_independent_of_threads_and_batch(small, tmp_path):
    want = H.expected(small, "out_bw5.sam.gz")
    for k, extra in enumerate((["-t", "4"], ["-batch", "37"], ["-t", "3", "-batch", "100"])):
        out = str(tmp_path / f"o{k}.sam")
        cmd = [HOST, "-x", small.idx_path, "-q", os.path.join(small.dir, "reads.fa"), "-osh", out] + extra
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        assert H.sam_lines(open(out).read()) == want


def test_sam_same_with_clumps_on_host_or_device(small, tmp_path):
    """Row N1 switches: fragments -> clumps and the first phase of their alignment on the device (default), the
    alignment phase by the workers (YA_HOST_PREP=1), or both by the workers (YA_HOST_CLUMPS=1)."""
    want = H.expected(small, "out_bw10.sam.gz")
    for k, env in enumerate(({}, {"YA_FUSED": "0"}, {"YA_HOST_PREP": "1"}, {"YA_HOST_CLUMPS": "1"})):
        out = str(tmp_path / f"c{k}.sam")
        cmd = H.command(HOST, small, "reads.fa", "-osh", out, ["-BW", "10", "-G", "100"], threads=3)
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
        assert p.returncode == 0, p.stderr[-2000:]
        assert H.sam_lines(open(out).read()) == want, env


def test_most_reads_are_finished_on_the_device(small, tmp_path):
    """Rows N2 / N4: with ya_align_batch (default) the reads of the golden set whose clumps need no split are aligned, scored,
    run through OQC and formatted on the device -- the host program's counters say how many; YA_FUSED=0 takes every read
    through the fibers.  Same SAM either way (and the reference's)."""
    import json
    want = H.expected(small, "out_multi_fbs.sam.gz")
    seen = {}
    for tag, env in (("fused", {}), ("fibers", {"YA_FUSED": "0"})):
        out = str(tmp_path / f"{tag}.sam")
        cmd = H.command(HOST, small, "multi.fa", "-osh", out, ["-FBS", "Y", "-PRL", "0.5", "-PSS", "0.5"], threads=2)
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, YAHA_B200_STATS="1", **env))
        assert p.returncode == 0, p.stderr[-2000:]
        assert H.sam_lines(open(out).read()) == want, tag
        seen[tag] = [json.loads(l) for l in p.stderr.splitlines() if l.startswith('{"pass"')][-1]
    assert seen["fused"]["reads_finished_on_device"] >= 200 and seen["fused"]["reads_handed_back"] <= 20
    assert seen["fused"]["device_text_bytes"] > 100_000
    assert seen["fibers"]["reads_finished_on_device"] == 0


def test_index_mode_writes_the_reference_files(small, tmp_path):
    """`yaha_b200_host -g ref.fa -L 11 -S 1`: .nib2 and index file (built on the device) under the reference's names with the
    digests of the files `yaha -g` wrote (golden/small/files.sha256); and one -S / -H variant against index_variants.json."""
    import hashlib
    import json
    import shutil
    shutil.copy(os.path.join(small.dir, "ref.fa"), tmp_path / "ref.fa")
    p = subprocess.run([HOST, "-g", str(tmp_path / "ref.fa"), "-L", "11"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    for name in ("ref.nib2", "ref.X11_01_65525S"):
        assert hashlib.sha256(open(tmp_path / name, "rb").read()).hexdigest() == small.sha[name], name
    want = json.load(open(os.path.join(small.golden, "index_variants.json")))
    shutil.copy(tmp_path / "ref.nib2", tmp_path / "small.nib2")
    p = subprocess.run([HOST, "-g", str(tmp_path / "small.nib2"), "-L", "9", "-S", "1", "-H", "20"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert hashlib.sha256(open(tmp_path / "small.X09_01_00020S", "rb").read()).hexdigest() == want["small.X09_01_00020S"]


def test_sam_identical_on_10k_long_reads(tmp_path):
    """BASELINE configs[0] shape at reduced reference size, against the reference binary when it is
    on this box (oracle/_ref travels with the snapshot): 1 Mbp, 2 000 x 1000 bp reads at 2 %."""
    if not os.path.exists(S.REF_BIN):
        pytest.skip("oracle/_ref/yaha not present")
    import numpy as np
    import yaha_b200
    from yaha_b200 import refio, synth
    d = str(tmp_path)
    ref = synth.random_reference(1_000_000, 31)
    synth.write_fasta(d + "/ref.fa", [("chrA", ref[:600_000]), ("chrB", ref[600_000:])])
    synth.write_reads(d + "/reads.fa", synth.simulate_reads(ref, 2000, 1000, 0.02, 5))
    open(d + "/ref.nib2", "wb").write(refio.build_nib2(refio.read_fasta(d + "/ref.fa")))
    nib = refio.load_nib2(d + "/ref.nib2")
    al = yaha_b200.Aligner(nib, None, yaha_b200.Params.defaults(word_len=13), device=0)
    idx_path = d + "/" + refio.index_file_name("ref", 13, 1, 65525)
    refio.write_index(idx_path, al.download_index())
    al.close()
    subprocess.run([S.REF_BIN, "-x", idx_path, "-q", d + "/reads.fa", "-osh", d + "/r.sam", "-t", "1"], check=True, capture_output=True)
    p = subprocess.run([HOST, "-x", idx_path, "-q", d + "/reads.fa", "-osh", d + "/m.sam", "-t", "4"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert H.sam_lines(open(d + "/m.sam").read()) == H.sam_lines(open(d + "/r.sam").read())


def test_reference_host_through_the_abi(small, tmp_path):
    """Mode A: the UNMODIFIED reference host, its hot-path seams redirected with
    -Wl,--wrap to libyaha_b200.so (oracle/modeA_shim.c), must write the stock binary's SAM."""
    modea = os.path.join(S.ORACLE_DIR, "_ref", "yaha_modeA")
    if not os.path.exists(modea):
        pytest.skip("oracle/_ref/yaha_modeA not present")
    out = str(tmp_path / "a.sam")
    p = subprocess.run([modea, "-x", small.idx_path, "-q", os.path.join(small.dir, "reads.fa"), "-osh", out, "-t", "1"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert H.sam_lines(open(out).read()) == H.expected(small, "out_bw5.sam.gz")
