"""Host logic on the CPU: the product's host sources (yaha_b200/host/*.cpp: fragment graph, clump
assembly, split/score, OQC/FBS, SAM writer, fiber scheduler) linked against an oracle-backed mock
of the C ABI (tests/mock/mock_abi.c) must reproduce the reference's SAM byte for byte.  The same
cases run against the real CUDA library in test_host_sam.py (-m gpu)."""
import os
import subprocess

import pytest

import hostcases as H
import support as S

MOCK = os.path.join(S.ROOT, "tests", "_build", "yaha_host_mock")


@pytest.fixture(scope="module")
def mock_host():
    subprocess.check_call(["make", "-s", "-C", os.path.join(S.ROOT, "tests", "mock"), "SAN="])
    return MOCK


@pytest.mark.parametrize("golden,reads,outflag,extra", H.CASES)
def test_host_logic_reproduces_reference_sam(small, mock_host, tmp_path, golden, reads, outflag, extra):
    out = str(tmp_path / "o.sam")
    p = subprocess.run(H.command(mock_host, small, reads, outflag, out, extra), capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    got, want = H.sam_lines(open(out).read()), H.expected(small, golden)
    diff = [(i, x, y) for i, (x, y) in enumerate(zip(got, want)) if x != y]
    assert not diff, diff[:2]
    assert len(got) == len(want)


@pytest.mark.parametrize("flags", H.FLAG_SWEEP, ids=lambda f: "".join(f))
def test_every_alignment_flag_reproduces_reference_sam(small, mock_host, tmp_path, flags):
    # -X -M -MD -P -H -AGS -GOC -GEC -MS -RC -BP -MGDP -MNO -PRL -PSS -BW -G, alone and combined, against digests of the
    # reference's output (golden/small/flag_sweep.json)
    H.check_flag_sweep(mock_host, small, str(tmp_path / "o.sam"), flags)


def test_output_order_independent_of_threads_and_batch(small, mock_host, tmp_path):
    want = H.expected(small, "out_bw5.sam.gz")
    for k, extra in enumerate((["-t", "3"], ["-batch", "37"], ["-t", "2", "-batch", "100"], ["-t", "4", "-gpus", "2", "-batch", "50", "-pipes", "2"],
                               ["-t", "3", "-pipes", "3", "-batch", "64"])):
        out = str(tmp_path / f"o{k}.sam")
        cmd = [mock_host, "-x", small.idx_path, "-q", os.path.join(small.dir, "reads.fa"), "-osh", out] + extra
        subprocess.run(cmd, check=True, capture_output=True, timeout=600)
        assert H.sam_lines(open(out).read()) == want


def test_sliced_and_sequential_fasta_readers_agree(small, mock_host, tmp_path):
    # FASTA is cut into records by the reader and parsed by the pipelines (RecordSlicer / parseFastaRecord);
    # YA_SEQ_READER=1 keeps the sequential reader of the FASTQ path.  Both must give the reference's records,
    # including the odd inputs (CRLF, over-long, too-short, empty and multi-line records, '>' inside a line).
    for reads, golden in (("weird.fa", "out_weird.sam.gz"), ("reads.fa", "out_bw5.sam.gz"), ("weird3.fa", "out_weird3.sam.gz")):
        want = H.expected(small, golden)
        extra = ["-BW", "5", "-G", "50"] if reads == "reads.fa" else []
        for k, env in enumerate(({}, {"YA_SEQ_READER": "1"})):
            out = str(tmp_path / f"r{k}.sam")
            cmd = [mock_host, "-x", small.idx_path, "-q", os.path.join(small.dir, reads), "-osh", out, "-t", "2", "-batch", "41"] + extra
            subprocess.run(cmd, check=True, capture_output=True, timeout=600, env=dict(os.environ, **env))
            assert H.sam_lines(open(out).read()) == want, (reads, env)


def test_clumps_and_alignment_phase_one_on_device_or_host(small, mock_host, tmp_path):
    # the mock runs csrc/form_clumps.h and csrc/prepare_clumps.h behind ya_form_clumps / ya_prepare_clumps, the very
    # sources the device kernels compile; the switches move either step back into the worker threads
    want = H.expected(small, "out_bw10.sam.gz")
    for k, env in enumerate(({}, {"YA_HOST_PREP": "1"}, {"YA_HOST_CLUMPS": "1"})):
        out = str(tmp_path / f"d{k}.sam")
        cmd = H.command(mock_host, small, "reads.fa", "-osh", out, ["-BW", "10", "-G", "100"], threads=3)
        subprocess.run(cmd, check=True, capture_output=True, timeout=600, env=dict(os.environ, **env))
        assert H.sam_lines(open(out).read()) == want, env


def test_cli_errors_match_reference_behaviour(small, mock_host):
    # missing -x for query mode, bad flag, bad bool: message + non-zero exit like Main.c
    p = subprocess.run([mock_host, "-q", "x.fa"], capture_output=True, text=True)
    assert p.returncode != 0 and "Index file specification (-x) is required" in p.stderr
    p = subprocess.run([mock_host, "-bogus"], capture_output=True, text=True)
    assert p.returncode != 0 and "is not a valid option" in p.stderr
    p = subprocess.run([mock_host, "-x", small.idx_path, "-q", "r.fa", "-OQC", "maybe"], capture_output=True, text=True)
    assert p.returncode != 0 and "is not a valid value for parameter -OQC" in p.stderr


@pytest.mark.parametrize("word_len,skip", [(12, 3), (13, 2), (10, 1)])
def test_other_word_lengths_and_skip_distances(small, mock_host, tmp_path, word_len, skip):
    # -L / -S are index-time flags (Main.c:559-563 puts them into the file name): the index is made by the unmodified
    # reference where it is built (oracle/_ref/yaha; skipped elsewhere), both programs align against that file
    if not os.path.exists(S.REF_BIN):
        pytest.skip("oracle/_ref/yaha not built here")
    import shutil
    shutil.copy(os.path.join(small.dir, "ref.fa"), tmp_path / "ref.fa")
    subprocess.run([S.REF_BIN, "-g", "ref.fa", "-L", str(word_len), "-S", str(skip)], cwd=tmp_path, check=True, capture_output=True, timeout=600)
    idx = str(tmp_path / f"ref.X{word_len:02d}_{skip:02d}_65525S")
    reads = os.path.join(small.dir, "reads.fa")
    subprocess.run([S.REF_BIN, "-x", idx, "-q", reads, "-osh", str(tmp_path / "want.sam"), "-t", "1"], check=True, capture_output=True, timeout=600)
    subprocess.run([mock_host, "-x", idx, "-q", reads, "-osh", str(tmp_path / "got.sam"), "-t", "2"], check=True, capture_output=True, timeout=600)
    assert H.sam_lines(open(tmp_path / "got.sam").read()) == H.sam_lines(open(tmp_path / "want.sam").read())


def test_queries_from_standard_input(small, mock_host, tmp_path):
    """Without -q the queries come from standard input, as in the reference (Main.c:173-178, Query.c:63-74): FASTA and FASTQ,
    same SAM as from the file.  (`-q stdin` fails in the reference -- it maps the name to "stdout" -- and here.)"""
    for golden, reads, outflag, extra in (("out_bw5.sam.gz", "reads.fa", "-osh", ["-BW", "5", "-G", "50"]), ("out_fastq_oss.sam.gz", "reads.fq", "-oss", [])):
        out = str(tmp_path / f"{reads}.sam")
        cmd = [mock_host, "-x", small.idx_path, outflag, out, "-t", "2"] + extra
        with open(os.path.join(small.dir, reads), "rb") as f:
            p = subprocess.run(cmd, stdin=f, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        assert H.sam_lines(open(out).read()) == H.expected(small, golden)


def test_index_mode_compresses_the_genome_like_the_reference(small, mock_host, tmp_path):
    """`-g genome.fa`: the .nib2 the C++ host writes (host/indexer.cpp; compressFile, Compress.c:140-331) has the digest of the
    file the unmodified reference wrote (golden/small/files.sha256).  The index itself is built on the device (ya_open_build):
    the mock has none and says so -- tests/test_host_sam.py checks both files on the GPU."""
    import hashlib
    import shutil
    shutil.copy(os.path.join(small.dir, "ref.fa"), tmp_path / "ref.fa")
    p = subprocess.run([mock_host, "-g", str(tmp_path / "ref.fa"), "-L", "11"], capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "cannot build the index" in p.stderr
    assert hashlib.sha256(open(tmp_path / "ref.nib2", "rb").read()).hexdigest() == small.sha["ref.nib2"]


def test_batch_that_does_not_fit_one_device_pass_takes_the_call_by_call_path(small, mock_host, tmp_path):
    """ya_align_batch answering YA_E_STATE (a batch with more seed hits than one pass holds) is not fatal: the host builds the
    reads from its flat input buffers and runs the whole batch through the fibers.  Same SAM."""
    out = str(tmp_path / "o.sam")
    p = subprocess.run(H.command(mock_host, small, "reads.fa", "-osh", out, ["-BW", "5", "-G", "50"], threads=2), capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, YA_MOCK_ALIGN_TOO_BIG="1"))
    assert p.returncode == 0, p.stderr[-2000:]
    assert H.sam_lines(open(out).read()) == H.expected(small, "out_bw5.sam.gz")
