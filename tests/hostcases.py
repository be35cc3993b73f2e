"""The (reference output, command line) pairs shared by the CPU mock-host test and the GPU host
test.  Every expected file was written by the unmodified reference (tests/golden/make_golden.py)."""
import gzip
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
]


def sam_lines(text):
    """All lines except @PG, which echoes file names and -t (AlignOutput.c:50-60)."""
    return [l for l in text.splitlines() if not l.startswith("@PG")]


def expected(small, golden):
    return sam_lines(gzip.open(os.path.join(small.golden, golden), "rt").read())


def command(binary, small, reads, outflag, out, extra, threads=1):
    return [binary, "-x", small.idx_path, "-q", os.path.join(small.dir, reads), outflag, out, "-t", str(threads)] + extra
