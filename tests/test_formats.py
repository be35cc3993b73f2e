"""On-disk formats stay bit-compatible with the reference."""
import numpy as np

from yaha_b200 import refio


def test_nib2_and_index_match_reference_digests(small):
    # digests were taken from files written by the unmodified reference (`yaha -g`)
    assert small.nib_sha == small.sha["ref.nib2"]
    assert small.idx_sha == small.sha["ref.X11_01_65525S"]


def test_nib2_roundtrip(small):
    nib = small.nib
    assert nib.names == ["chr1", "chr2", "chr3"]
    assert list(nib.lengths) == [120000, 80003, 99997]
    # sequences start on 4-byte (8-base) boundaries, padded with X (14)
    assert all(int(s) % 8 == 0 for s in nib.starts)
    assert nib.max_roff == int(nib.starts[-1] + nib.lengths[-1])
    seqs = refio.read_fasta(small.dir + "/ref.fa")
    for (name, s), st, ln in zip(seqs, nib.starts, nib.lengths):
        assert np.array_equal(nib.unpack(int(st), int(ln)), refio.encode(s))
    pad = nib.unpack(int(nib.starts[1] + nib.lengths[1]), 5)
    assert list(pad) == [14] * 5


def test_index_invariants(small):
    idx = small.idx
    assert idx.word_len == 11 and idx.max_hits == 65525
    so = np.asarray(idx.so, dtype=np.int64)
    assert so[0] == 0 and so[-1] == idx.total and np.all(np.diff(so) >= 0)
    roa = np.asarray(idx.roa)
    # each k-mer list ascending; spot check a few hundred non-empty ones
    ne = np.nonzero(np.diff(so) > 1)[0][:300]
    for h in ne:
        lst = roa[so[h]:so[h + 1]]
        assert np.all(np.diff(lst.astype(np.int64)) > 0)


def test_code_tables():
    # Math.c:141-156
    assert list(refio.encode(b"TCAGNBDHKMRSVWXYtcagnU*")) == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15,
                                                               0, 1, 2, 3, 4, 0, 14]
    assert list(refio.COMP_CODE) == [2, 3, 0, 1, 4, 12, 7, 6, 9, 8, 15, 11, 5, 13, 14, 10]


def test_downsample_matches_reference_rng():
    # Marsaglia xorshift, default seed (Math.c:256-284): first outputs are fixed
    st = [123456789, 362436069, 521288629, 88675123, 886756453]
    a = [refio._marsaglia(st) for _ in range(3)]
    assert all(0 <= x < 2 ** 32 for x in a) and len(set(a)) == 3
    inp = np.arange(100, dtype=np.uint32)
    st = [123456789, 362436069, 521288629, 88675123, 886756453]
    out = refio._rand_sample(st, inp, 10)
    assert len(out) == 10 and np.all(np.diff(out.astype(int)) > 0)
    st = [123456789, 362436069, 521288629, 88675123, 886756453]
    out = refio._rand_sample(st, inp, 90)
    assert len(out) == 90 and np.all(np.diff(out.astype(int)) > 0)
