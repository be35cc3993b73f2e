"""The CPU oracle (oracle/*.c) is pinned against the unmodified reference: every call the
reference made through the hot-path seams while aligning the golden reads (recorded by
oracle/dump_shim.c into tests/golden/small/dump_*.txt.gz) must be reproduced exactly."""
import numpy as np
import pytest

import support as S


def _params(small, bw, gap):
    return S.default_params(word_len=small.idx.word_len, max_hits=min(650, small.idx.max_hits), bw=bw, max_gap=gap)


@pytest.mark.parametrize("bw,gap", [(5, 50), (10, 100)])
def test_dp_jobs_match_reference(small, bw, gap):
    p = _params(small, bw, gap)
    n = 0
    kinds = set()
    for rec in S.parse_dump(small.dump(bw), "D"):
        _, k, qid, st, roff, rlen, qoff, qlen, score, aq, ar, ops = rec
        got = S.oracle_dp(p, small.nib.bases, small.nib.max_roff, small.codes(qid, st), S.KIND_OF_CHAR[k],
                          roff, rlen, qoff, qlen)
        assert got[:4] == (score, aq, ar, ops), (rec, got)
        kinds.add(k)
        n += 1
    assert n > 3000 and kinds == {"F", "B", "E", "R"}


def test_seed_lookup_and_fragments_match_reference(small):
    p = _params(small, 5, 50)
    frags_of = {}
    n_s = n_g = n_quirk = 0
    for rec in S.parse_dump(small.dump(5), "SG"):
        if rec[0] == "S":
            _, qid, st, mc, ent = rec
            soff, cnt, total, frags, region, keep = S.oracle_seed_frags(p, small.idx.so, small.idx.roa, small.codes(qid, st))
            assert len(cnt) == mc
            assert [(int(i), int(soff[i]), int(cnt[i])) for i in np.nonzero(cnt)[0]] == ent
            frags_of[(qid, st)] = (frags, total)
            n_s += 1
        else:
            _, qid, st, fr = rec
            frags, total = frags_of.pop((qid, st))
            mine = [(int(f["startRefOff"]), int(f["startQueryOff"]), int(f["endQueryOff"]), int(f["refLen"])) for f in frags]
            assert mine == fr, (qid, st)
            # the over-read quirk of QueryMatch.c:62-67 makes more hits than totalCount
            if sum(1 for _ in fr) and qid.startswith("head"):
                n_quirk += 1
            n_g += 1
    assert n_s == n_g and n_s > 1000 and n_quirk > 0


def test_singleton_clumps_match_reference(small):
    """Stage 2b: regions with one fragment become clumps iff refLen >= minMatch
    (QueryMatch.c:281-290).  The C records list the clump list after each strand; single-fragment
    clumps that the graph did not make must equal the oracle's surviving singletons."""
    p = _params(small, 5, 50)
    seen = 0
    prev_clumps = {}
    for rec in S.parse_dump(small.dump(5), "GC"):
        if rec[0] == "G":
            _, qid, st, fr = rec
            cur = (qid, st, fr)
            continue
        _, qid, st, clumps = rec
        assert cur[0] == qid and cur[1] == st
        fr = cur[2]
        frags = np.zeros(len(fr), dtype=S.FRAG_DT)
        for i, (sro, sqo, eqo, rl) in enumerate(fr):
            frags[i] = (sro, sqo, eqo, 0, rl)
        region = np.zeros(len(fr), dtype=np.uint32)
        keep = np.zeros(len(fr), dtype=np.uint8)
        if len(fr):
            S.oracle().orc_regions(p, S.ptr(frags), len(fr), S.ptr(region), S.ptr(keep))
        members = np.bincount(region, minlength=1) if len(fr) else np.zeros(0, int)
        want = [fr[i] for i in range(len(fr)) if members[region[i]] == 1 and keep[i]]
        # clumps pushed during this strand = new head entries (LIFO, QueryState.c:156-161)
        before = prev_clumps.get(qid, 0) if st == 1 else 0
        new = clumps[:len(clumps) - before]
        prev_clumps[qid] = len(clumps)
        got_single = [c[1][0] for c in new if len(c[1]) == 1 and c[0] == st]
        for w in want:
            assert w in got_single, (qid, st, w)
        seen += len(want)
    assert seen >= 3


@pytest.mark.parametrize("bw,gap", [(5, 50)])
def test_multi_fragment_clumps_match_reference(small, bw, gap):
    """Row N1 pinned independently of the SAM: the CPU build of csrc/form_clumps.h (orc_form_clumps -- the source the
    device kernel and the host program compile) must reproduce EVERY clump the unmodified reference formed on the
    golden reads, multi-fragment ones included: same clumps in the same creation order, same fragments after the
    overlap chops of insertFragment and cleanUpClump (QueryMatch.c:224-303, GraphPath.cpp:161-292,
    AlignHelpers.c:48-193).  The C records of the dump list the read's clump list (head first, LIFO:
    QueryState.c:156-161) after each strand's processFragmentsGapped."""
    p = _params(small, bw, gap)
    lib = S.oracle()
    n_multi = n_clumps = 0
    before = {}
    cur = None
    for rec in S.parse_dump(small.dump(bw), "GC"):
        if rec[0] == "G":
            cur = rec
            continue
        _, qid, st, clumps = rec
        assert cur[1] == qid and cur[2] == st
        fr = cur[3]
        n_before = before.get(qid, 0) if st == 1 else 0
        new = clumps[:len(clumps) - n_before][::-1]                   # creation order
        before[qid] = len(clumps)
        frags = np.zeros(len(fr), dtype=S.FRAG_DT)
        for i, (sro, sqo, eqo, rl) in enumerate(fr):
            frags[i] = (sro, sqo, eqo, 0, rl)
        got = []
        if len(fr):
            region = np.zeros(len(fr), dtype=np.uint32)
            keep = np.zeros(len(fr), dtype=np.uint8)
            lib.orc_regions(p, S.ptr(frags), len(fr), S.ptr(region), S.ptr(keep))
            sel = keep.astype(bool)
            sf, sr = np.ascontiguousarray(frags[sel]), np.ascontiguousarray(region[sel])
            n = len(sf)
            if n:
                L = len(small.codes(qid, st))
                opath = np.zeros(n, dtype=S.FRAG_DT)
                ocl = np.zeros(n, dtype=np.dtype([("first", "<u4"), ("n", "<u2"), ("matchedBases", "<u2")]))
                nc = lib.orc_form_clumps(p.wordLen, gap, 50, 25, 25, bw, 5, 2, 1, S.ptr(sf), S.ptr(sr), n, L, S.ptr(opath), S.ptr(ocl))
                for k in range(nc):
                    a = opath[int(ocl[k]["first"]):int(ocl[k]["first"]) + int(ocl[k]["n"])]
                    got.append([(int(f["startRefOff"]), int(f["startQueryOff"]), int(f["endQueryOff"]), int(f["refLen"])) for f in a])
        want = [list(c[1]) for c in new]
        assert all(c[0] == st for c in new)
        assert got == want, (qid, st, got, want)
        n_clumps += len(want)
        n_multi += sum(1 for w in want if len(w) > 1)
    assert n_multi > 500 and n_clumps > n_multi
