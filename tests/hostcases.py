"""The (reference output, command line) pairs shared by the CPU mock-host test and the GPU host
test.  Every expected file was written by the unmodified reference (tests/golden/make_golden.py)."""
import gzip
import hashlib
import json
import os

CASES = [
    # golden file,        reads,      output flag, extra flags
    ("out_bw5.sam.gz", "reads.fa", "-osh", ["-BW", "5", "-G", "50"]),
    ("out_bw10.sam.gz", "reads.fa", "-osh", ["-BW", "10", "-G", "100"]),
    ("out_fbs.sam.gz", "reads.fa", "-osh", ["-FBS", "Y"]),
    ("out_nooqc.sam.gz", "reads.fa", "-osh", ["-OQC", "N"]),
    ("out_fastq_oss.sam.gz", "reads.fq", "-oss", []),
    ("out_blast8.sam.gz", "reads.fa", "-o8", []),
    ("out_weird.sam.gz", "weird.fa", "-osh", []),
    # reads glued from 2-6 loci whose pieces each have to be split: several splitting clumps per read (child fibers)
    ("out_chimera.sam.gz", "chimera.fa", "-osh", []),
    # FASTQ reader corner cases: descriptions, CRLF, multi-line records, '@' inside quality, mismatched lengths, no final newline
    ("out_weird_fastq.sam.gz", "weird.fq", "-oss", []),
    # FASTA id lines holding the record markers themselves ('>' '+' '@'): an id line ends at its newline only
    ("out_weird2.sam.gz", "weird2.fa", "-osh", []),
    # reads glued from 2-7 clean pieces (overlapping, duplicated, subsumed, equal-scoring): Optimal Query Coverage, the filter
    # by similarity and the mapping qualities on reads the device finishes itself (csrc/finish_reads.h)
    ("out_multi.sam.gz", "multi.fa", "-osh", []),
    ("out_multi_fbs.sam.gz", "multi.fa", "-osh", ["-FBS", "Y", "-PRL", "0.5", "-PSS", "0.5"]),
    ("out_multi_mno.sam.gz", "multi.fa", "-osh", ["-MNO", "5", "-BP", "2", "-MGDP", "9", "-M", "15"]),
    # an empty record in front (behind one skipped as too short) is read over once (Query.c:304,637); the next empty record ends the run
    ("out_weird3.sam.gz", "weird3.fa", "-osh", []),
    ("out_weird3_fastq.sam.gz", "weird3.fq", "-oss", []),
]

# One line per flag of the reference's alignment CLI (Main.c:187-470) that changes the result, plus combinations and
# degenerate values (-X 0, -BW 0, -H 1, band wider than the gap cap).  All on reads.fa with -osh; the reference's output for each
# set is pinned as line count + sha256 of its non-@PG lines in golden/small/flag_sweep.json (make_golden.py --only-flag-sweep).
FLAG_SWEEP = [
    ["-X", "10"], ["-X", "60"], ["-X", "0"],
    ["-M", "15"], ["-M", "40"], ["-M", "12"],
    ["-MD", "10"], ["-MD", "100"],
    ["-P", "0.8"], ["-P", "0.99"], ["-P", "0.5"],
    ["-H", "5"], ["-H", "1"], ["-H", "10000"],
    ["-AGS", "N"], ["-AGS", "N", "-BW", "8", "-G", "80"],
    ["-GOC", "0"], ["-GOC", "10", "-GEC", "1"], ["-GEC", "5"], ["-MS", "2"],
    ["-MS", "3", "-RC", "5", "-GOC", "8", "-GEC", "3"], ["-RC", "1"], ["-RC", "10"],
    ["-BP", "1"], ["-BP", "20"], ["-MGDP", "1"], ["-MGDP", "9"], ["-MNO", "1"], ["-MNO", "60"],
    ["-FBS", "Y", "-PRL", "0.5"], ["-FBS", "Y", "-PSS", "0.5"], ["-FBS", "Y", "-PRL", "1.0", "-PSS", "1.0"],
    ["-BW", "1", "-G", "10"], ["-BW", "20", "-G", "200"], ["-BW", "3", "-G", "500"], ["-BW", "30", "-G", "30"],
    ["-BW", "50", "-G", "120"], ["-BW", "0"], ["-G", "5"], ["-G", "1000"],
    ["-BW", "10", "-G", "100", "-X", "40", "-M", "20", "-MD", "30", "-P", "0.85", "-MS", "2", "-RC", "4", "-GOC", "6", "-GEC", "2"],
    ["-OQC", "N", "-M", "15", "-P", "0.7"],
]


def sam_lines(text):
    """All lines except @PG, which echoes file names and -t (AlignOutput.c:50-60)."""
    return [l for l in text.splitlines() if not l.startswith("@PG")]


def expected(small, golden):
    return sam_lines(gzip.open(os.path.join(small.golden, golden), "rt").read())


def command(binary, small, reads, outflag, out, extra, threads=1):
    return [binary, "-x", small.idx_path, "-q", os.path.join(small.dir, reads), outflag, out, "-t", str(threads)] + extra


def digest(lines):
    return {"lines": len(lines), "sha256": hashlib.sha256("\n".join(lines).encode()).hexdigest()}


def flag_sweep_expected(small):
    return json.load(open(os.path.join(small.golden, "flag_sweep.json")))


def check_flag_sweep(binary, small, out, flags, threads=2, env=None):
    """Runs `binary` with one FLAG_SWEEP set and compares the digest of its SAM with the reference's."""
    import subprocess
    p = subprocess.run(command(binary, small, "reads.fa", "-osh", out, flags, threads=threads), capture_output=True, text=True,
                       timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    want = flag_sweep_expected(small)[" ".join(flags)]
    assert digest(sam_lines(open(out).read())) == want, flags
