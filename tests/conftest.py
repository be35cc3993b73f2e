import gzip
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from yaha_b200 import refio  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class Small:
    """The committed golden fixture tests/golden/small (made by tests/golden/make_golden.py from
    the unmodified reference).  nib2 + index are rebuilt with our own writers and must hash to
    the digests of the reference-built files."""

    def __init__(self, tmpdir):
        g = os.path.join(ROOT, "tests", "golden", "small")
        self.dir = str(tmpdir)
        for name in ("ref.fa", "reads.fa", "reads.fq", "weird.fa", "chimera.fa", "weird.fq", "weird2.fa", "multi.fa", "weird3.fa", "weird3.fq"):
            with gzip.open(os.path.join(g, name + ".gz"), "rb") as f, open(os.path.join(self.dir, name), "wb") as o:
                o.write(f.read())
        self.golden = g
        self.sha = dict(line.split()[::-1] for line in open(os.path.join(g, "files.sha256")))
        nib_img = refio.build_nib2(refio.read_fasta(os.path.join(self.dir, "ref.fa")))
        self.nib_path = os.path.join(self.dir, "ref.nib2")
        open(self.nib_path, "wb").write(nib_img)
        self.nib_sha = hashlib.sha256(nib_img).hexdigest()
        self.nib = refio.load_nib2(self.nib_path)
        idx_img = refio.build_index(self.nib, 11)
        self.idx_path = os.path.join(self.dir, refio.index_file_name("ref", 11, 1, 65525))
        open(self.idx_path, "wb").write(idx_img)
        self.idx_sha = hashlib.sha256(idx_img).hexdigest()
        self.idx = refio.load_index(self.idx_path)
        self.reads = refio.read_queries(os.path.join(self.dir, "reads.fa"), word_len=11)
        self.names = [n for n, _ in self.reads]
        self.read_id = {n: i for i, n in enumerate(self.names)}
        self.fwd = [np.ascontiguousarray(refio.encode(s)) for _, s in self.reads]
        self.rev = [np.ascontiguousarray(refio.revcomp_codes(c)) for c in self.fwd]

    def dump(self, bw):
        return os.path.join(self.golden, f"dump_bw{bw}.txt.gz")

    def codes(self, name, strand):
        i = self.read_id[name]
        return self.rev[i] if strand else self.fwd[i]


@pytest.fixture(scope="session")
def small(tmp_path_factory):
    return Small(tmp_path_factory.mktemp("small"))
