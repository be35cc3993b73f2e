"""Full-size checks through the C ABI at BASELINE.json configs[2] (100 Mbp reference, 20K reads of
500 bp at 10 % error, -BW 10 -G 100), where the oracle is too slow to replay everything: the
size-independent properties of the three device stages, plus an oracle comparison on a seeded
sample.  (bench.py additionally diffs the whole SAM of this workload against the reference binary.)
"""
import os

import numpy as np
import pytest

import support as S
import yaha_b200
from yaha_b200 import refio, synth

pytestmark = pytest.mark.gpu

REF_BASES, N_READS, READ_LEN, ERR = 100_000_000, 20_000, 500, 0.10
BW, GAP = 10, 100


@pytest.fixture(scope="module")
def big(tmp_path_factory):
    d = tmp_path_factory.mktemp("full")
    ref = synth.random_reference(REF_BASES, 12345)
    path = os.path.join(str(d), "ref.nib2")
    with open(path, "wb") as f:
        f.write(refio.build_nib2([("chr1", ref)]))
    nib = refio.load_nib2(path)
    P = yaha_b200.Params.defaults(word_len=15, bw=BW, max_gap=GAP)
    al = yaha_b200.Aligner(nib, None, P, device=0)            # index built on the device
    reads = list(synth.simulate_reads(ref, N_READS, READ_LEN, ERR, 777))
    fwd = [np.ascontiguousarray(refio.encode(r)) for _, r in reads]
    rev = [np.ascontiguousarray(refio.revcomp_codes(c)) for c in fwd]
    truth = [(int(n.split("_")[1]), n.endswith("-")) for n, _ in reads]
    al.upload_read_list(fwd)
    strands, frags, region = al.seed_frags()
    yield dict(nib=nib, al=al, fwd=fwd, rev=rev, truth=truth, strands=strands, frags=frags, region=region, P=P)
    al.close()


def _strand_of_frag(strands):
    return np.repeat(np.arange(len(strands)), strands["n_frags"].astype(np.int64))


def test_fragments_full_size(big):
    strands, frags, region = big["strands"], big["frags"], big["region"]
    n = len(frags)
    assert n > 8 * N_READS
    assert int(strands["n_frags"].sum()) == n
    ne = strands["n_frags"] > 0
    assert np.array_equal(strands["first"].astype(np.int64)[ne], (np.cumsum(strands["n_frags"].astype(np.int64)) - strands["n_frags"])[ne])
    # a fragment is a maximal run of identical bases on one diagonal: refLen == query length
    qlen = frags["endQueryOff"].astype(np.int64) - frags["startQueryOff"] + 1
    assert np.array_equal(qlen, frags["refLen"].astype(np.int64))
    assert int(qlen.min()) >= 15
    # within a strand: sorted by (diagonal, query offset); region ids never decrease
    sid = _strand_of_frag(strands)
    diag = frags["startRefOff"].astype(np.int64) - frags["startQueryOff"]
    key = (sid << 48) + ((diag + 65536) << 16) + frags["startQueryOff"]
    assert np.all(np.diff(key) > 0)
    same = sid[1:] == sid[:-1]
    assert np.all(region[1:][same] >= region[:-1][same])
    # exact-match property on a seeded sample, straight from the packed reference
    rng = np.random.default_rng(5)
    for i in rng.integers(0, n, size=4000):
        f = frags[i]
        r, st = divmod(int(sid[i]), 2)
        codes = (big["rev"] if st else big["fwd"])[r]
        want = big["nib"].unpack(int(f["startRefOff"]), int(f["refLen"]))
        assert np.array_equal(want, codes[int(f["startQueryOff"]):int(f["endQueryOff"]) + 1]), i
    # recall: the true locus carries a fragment for (nearly) every read
    hit = 0
    for r, (start, minus) in enumerate(big["truth"]):
        s = strands[2 * r + (1 if minus else 0)]
        a, k = int(s["first"]), int(s["n_frags"])
        d = diag[a:a + k]
        hit += bool(np.any(np.abs(d - start) <= 60))
    assert hit >= 0.99 * N_READS, hit


def test_fragments_sample_equals_oracle(big):
    """Bit-exact against the oracle on a seeded sample of strands (the oracle gathers from the index
    downloaded from the device, which test_gpu_parity proves identical to the reference's)."""
    idx = big["al"].download_index()
    p = S.default_params(word_len=15, bw=BW, max_gap=GAP)
    rng = np.random.default_rng(6)
    for r in rng.integers(0, N_READS, size=150):
        for st in (0, 1):
            codes = (big["rev"] if st else big["fwd"])[r]
            _, _, total, of, oreg, keep = S.oracle_seed_frags(p, idx.so, idx.roa, codes)
            s = big["strands"][2 * r + st]
            k = keep.astype(bool)
            a, n = int(s["first"]), int(s["n_frags"])
            assert (int(s["total_hits"]), int(s["n_frags_all"]), n) == (total, len(of), int(k.sum()))
            for f in ("startRefOff", "startQueryOff", "endQueryOff", "refLen"):
                assert np.array_equal(big["frags"][a:a + n][f], of[k][f])
            assert np.array_equal(big["region"][a:a + n], oreg[k])


def test_seed_sharding_is_linear_full_size(big):
    """Halves of the batch give exactly the slices of the whole (what the multi-GPU sharding relies on)."""
    al2 = yaha_b200.Aligner(big["nib"], None, big["P"], device=0)
    half = N_READS // 2
    base = 0
    for lo, hi in ((0, half), (half, N_READS)):
        al2.upload_read_list(big["fwd"][lo:hi])
        st, fr, rg = al2.seed_frags()
        whole = big["strands"][2 * lo:2 * hi]
        assert np.array_equal(st["n_frags"], whole["n_frags"])
        assert np.array_equal(st["total_hits"], whole["total_hits"])
        k = len(fr)
        assert fr.tobytes() == big["frags"][base:base + k].tobytes()
        base += k
    assert base == len(big["frags"])
    al2.close()


def _jobs_from_fragments(big):
    """Extension jobs off both ends of the longest true-locus fragment of every read, and gap-fill
    jobs between neighbouring fragments of a region (the job mix the host posts, AlignHelpers.c:205-290)."""
    strands, frags, region = big["strands"], big["frags"], big["region"]
    jobs = []
    for r, (start, minus) in enumerate(big["truth"]):
        st = 1 if minus else 0
        s = strands[2 * r + st]
        a, k = int(s["first"]), int(s["n_frags"])
        if k == 0:
            continue
        f = frags[a:a + k]
        b = int(np.argmax(f["refLen"]))
        sro, sqo, eqo, rl = (int(f[b][x]) for x in ("startRefOff", "startQueryOff", "endQueryOff", "refLen"))
        if eqo + 1 < READ_LEN + 60 and eqo + 1 < len(big["fwd"][r]):
            jobs.append((sro + rl, r, 0, eqo + 1, len(big["fwd"][r]) - eqo - 1, yaha_b200.DP_EXT_FWD, st))
        if sqo > 0 and sro > 0:
            jobs.append((sro - 1, r, 0, sqo - 1, sqo, yaha_b200.DP_EXT_BWD, st))
        order = np.argsort(f["startQueryOff"], kind="stable")
        for u, v in zip(order[:-1], order[1:]):
            if region[a + u] != region[a + v]:
                continue
            q0, q1 = int(f[u]["endQueryOff"]) + 1, int(f[v]["startQueryOff"])
            r0, r1 = int(f[u]["startRefOff"]) + int(f[u]["refLen"]), int(f[v]["startRefOff"])
            ql, rl2 = q1 - q0, r1 - r0
            if ql < 1 or rl2 < 1 or ql > 200 or abs(ql - rl2) > GAP:
                continue
            kind = yaha_b200.DP_BANDED if min(ql, rl2) > 2 * BW and abs(ql - rl2) <= BW else yaha_b200.DP_FULL
            jobs.append((r0, r, rl2, q0, ql, kind, st))
    return np.array(jobs, dtype=yaha_b200.JOB_DT)


def test_dp_full_size(big):
    al, P = big["al"], big["P"]
    jobs = _jobs_from_fragments(big)
    assert len(jobs) > 60_000
    res, ops = al.sw_batch(jobs)
    # determinism and linearity: the same jobs again, and in two halves, give the same bytes
    res2, ops2 = al.sw_batch(jobs)
    assert res.tobytes() == res2.tobytes() and ops.tobytes() == ops2.tobytes()
    h = len(jobs) // 2
    ra, oa = al.sw_batch(jobs[:h])
    rb, ob = al.sw_batch(jobs[h:])
    for f in ("score", "addedQLen", "addedRLen", "ops_n"):
        assert np.array_equal(np.concatenate([ra[f], rb[f]]), res[f]), f
    assert np.concatenate([oa, ob]).tobytes() == ops.tobytes()

    # edit scripts are self-consistent for EVERY job: lengths consumed, affine score, no equal neighbours
    n = len(jobs)
    off, cnt = res["ops_off"].astype(np.int64), res["ops_n"].astype(np.int64)
    assert np.array_equal(off, np.cumsum(cnt) - cnt) and int(cnt.sum()) == len(ops)
    owner = np.repeat(np.arange(n), cnt)
    ln = ops["length"].astype(np.int64)
    code = ops["opcode"]
    isM, isR, isI, isD = (code == ord(c) for c in "MRID")
    assert np.all(isM | isR | isI | isD) and np.all(ln > 0)

    def per_job(mask, w=None):
        return np.bincount(owner[mask], weights=(ln if w is None else w)[mask], minlength=n).astype(np.int64)
    m, r_, i_, d_ = per_job(isM), per_job(isR), per_job(isI), per_job(isD)
    glob = jobs["kind"] <= yaha_b200.DP_BANDED
    ext = ~glob
    # extensions report what they added; global jobs consume exactly their rectangle (added* stay 0)
    assert np.array_equal((m + r_ + i_)[ext], res["addedQLen"].astype(np.int64)[ext])
    assert np.array_equal((m + r_ + d_)[ext], res["addedRLen"].astype(np.int64)[ext])
    assert np.array_equal((m + r_ + i_)[glob], jobs["qLen"].astype(np.int64)[glob])
    assert np.array_equal((m + r_ + d_)[glob], jobs["rLen"].astype(np.int64)[glob])
    gaps = np.bincount(owner[isI | isD], minlength=n)
    score = P.MScore * m - P.RCost * r_ - P.GOCost * gaps - P.GECost * (i_ + d_)
    assert np.array_equal(score, res["score"].astype(np.int64))
    assert np.all(res["score"][ext] >= 0)
    assert np.all(res["addedQLen"][ext].astype(np.int64) <= jobs["qLen"][ext])
    sameowner = owner[1:] == owner[:-1]
    assert not np.any(sameowner & (code[1:] == code[:-1]))

    # M / R labels replayed against the sequences, and the oracle, on a seeded sample
    rng = np.random.default_rng(9)
    p = S.default_params(word_len=15, bw=BW, max_gap=GAP)
    nib = big["nib"]
    for i in rng.integers(0, n, size=2500):
        j = jobs[i]
        codes = (big["rev"] if j["strand"] else big["fwd"])[int(j["read"])]
        mine = ops[off[i]:off[i] + cnt[i]]
        kind, ro, qo = int(j["kind"]), int(j["rOff"]), int(j["qOff"])
        step = -1 if kind == yaha_b200.DP_EXT_BWD else 1
        for o in (mine[::-1] if step < 0 else mine):          # backward scripts are stored in read order
            L, c = int(o["length"]), chr(int(o["opcode"]))
            if c in "MR":
                for t in range(L):
                    eq = nib.base(ro + step * t) == int(codes[qo + step * t])
                    assert eq == (c == "M"), (i, tuple(j))
                ro += step * L
                qo += step * L
            elif c == "I":
                qo += step * L
            else:
                ro += step * L
        w = S.oracle_dp(p, nib.bases, nib.max_roff, codes, kind, int(j["rOff"]), int(j["rLen"]), int(j["qOff"]), int(j["qLen"]))[:4]
        got = (int(res[i]["score"]), int(res[i]["addedQLen"]), int(res[i]["addedRLen"]), S.ops_to_str(mine))
        assert got == tuple(w), (i, tuple(j))
