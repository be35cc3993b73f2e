"""csrc/assemble_clumps.h (what follows a clump's first DP round) against a step-by-step
restatement of the reference's own list operations.

The header writes a clump's runs front to back in one pass.  The reference (and this test's `expected`) builds the
list the long way: one list per piece ("nM" for a seed fragment, the gap answer between two of them), concatenated
with mergeEOLToBack (AlignHelpers.c:283-300, SW.cpp:207-261); the first and last run lengthened by the perfect end
extensions (AlignExtFrag.cpp:76-107); the backward extension's list merged in front (mergeEOLToFront, SW.cpp:151-205),
the forward extension's behind; then scoreClump's walk (AlignHelpers.c:302-366).  Inputs are seeded random clumps:
1-12 seed fragments, every gap kind (pure D, pure I, 1x1 R, DP answer with arbitrary runs and negative scores), clumps
hugging both ends of the reference and of the read, extensions accepted / refused / shorter than minExtLength.

The same routine is the checker for the device kernel of row N2 once it exists (one thread per clump runs the header).
"""
import ctypes as C

import numpy as np
import pytest

import support as S

GAP_DT = np.dtype([("job", "<u4"), ("score", "<i4"), ("after", "<u2"), ("len", "<u2"), ("code", "u1"), ("pad", "u1"), ("pad2", "<u2")])
PREP_DT = np.dtype([("gap_first", "<u4"), ("jobB", "<u4"), ("jobF", "<u4"), ("n_gaps", "<u2"), ("backLen", "<u2"),
                    ("forwLen", "<u2"), ("pad", "<u2")])
RES_DT = np.dtype([("score", "<i4"), ("addedQLen", "<u2"), ("addedRLen", "<u2"), ("ops_off", "<u4"), ("ops_n", "<u4")])
ASM_DT = np.dtype([("frag", S.FRAG_DT), ("score", "<i4"), ("n_ops", "<u4"), ("matchedBases", "<u2"), ("mismatchedBases", "<u2"),
                   ("gapBases", "<u2"), ("totLength", "<u2"), ("totScore", "<u2"), ("verdict", "u1"), ("pad", "u1"), ("ops_off", "<u4")])
NONE = 0xFFFFFFFF
DROP, SCORED, SPLIT = 0, 1, 2


def test_record_layouts_match_the_abi():
    assert (GAP_DT.itemsize, PREP_DT.itemsize, RES_DT.itemsize, S.OP_DT.itemsize, S.FRAG_DT.itemsize, ASM_DT.itemsize) == (16, 20, 16, 4, 12, 36)


def pack(codes):
    """4 bit per base, high nibble = even offset (Compress.c:251-283)."""
    c = np.concatenate([codes, np.zeros(len(codes) & 1, np.uint8)])
    return ((c[0::2] << 4) | c[1::2]).astype(np.uint8)


def merge_back(dst, src):                                   # mergeEOLToBack, SW.cpp:207-261
    src = [list(x) for x in src]
    if dst and src and dst[-1][0] == src[0][0]:
        dst[-1][1] = (dst[-1][1] + src[0][1]) & 0xFFFF
        src = src[1:]
    dst.extend(src)


def merge_front(dst, src):                                  # mergeEOLToFront, SW.cpp:151-205
    src = [list(x) for x in src]
    if dst and src and src[-1][0] == dst[0][0]:
        src[-1][1] = (src[-1][1] + dst[0][1]) & 0xFFFF
        del dst[0]
    dst[:0] = src


def op_score(P, code, ln):                                  # AlignHelpers.c:312-328
    if code == "M":
        return P["MS"] * ln
    if code == "R":
        return -(P["RC"] * ln)
    return -(P["GOC"] + P["GEC"] * ln)


def expected(P, genome, q, path, gaps, prep, answers, max_roff):
    """The reference's order of operations on Python lists.  Returns (ops, frag dict, score, verdict, counts)."""
    L = len(q)
    ops, score = [], 0
    by_after = {g["after"]: g for g in gaps}
    for it, f in enumerate(path):                           # collapseSFragments over seed pieces and gap pieces
        ql = 1 + f["eq"] - f["sq"]
        merge_back(ops, [["M", ql]])
        score += P["MS"] * ql
        g = by_after.get(it)
        if g is not None:
            if g["job"] is not None:
                a = answers[g["job"]]
                merge_back(ops, a["ops"])
                score += a["score"]
            else:
                merge_back(ops, [[g["code"], g["len"]]])
                score += g["score"]
    sq, sr = path[0]["sq"], path[0]["sr"]
    eq, er = path[-1]["eq"], path[-1]["sr"] + path[-1]["rl"] - 1
    # perfect end extensions (AlignExtFrag.cpp:30-48, 76-107)
    back = min(sq, sr)
    m = 0
    while m < back and q[sq - 1 - m] == genome[sr - 1 - m]:
        m += 1
    sq, sr, back = sq - m, sr - m, back - m
    ops[0][1] = (ops[0][1] + m) & 0xFFFF
    score += m * P["MS"]
    forw = min((L - 1) - eq, max_roff - er)
    m = 0
    while m < forw and q[eq + 1 + m] == genome[er + 1 + m]:
        m += 1
    eq, er, forw = eq + m, er + m, forw - m
    ops[-1][1] = (ops[-1][1] + m) & 0xFFFF
    score += m * P["MS"]
    assert (back >= P["minExt"]) == (prep["jobB"] is not None) and (forw >= P["minExt"]) == (prep["jobF"] is not None)
    if prep["jobB"] is not None:                            # AlignExtFrag.cpp:109-125
        a = answers[prep["jobB"]]
        if a["score"] > 0:
            merge_front(ops, a["ops"])
            score += a["score"]
            sq, sr = sq - a["aq"], sr - a["ar"]
    if prep["jobF"] is not None:                            # AlignExtFrag.cpp:127-143
        a = answers[prep["jobF"]]
        if a["score"] > 0:
            merge_back(ops, a["ops"])
            score += a["score"]
            eq, er = eq + a["aq"], er + a["ar"]
    # scoreClump, AlignHelpers.c:302-366
    ags = mx = 0
    cnt = dict(M=0, R=0, I=0, D=0)
    split = False
    for k, (code, ln) in enumerate(ops):
        cnt[code] += ln
        ags += op_score(P, code, ln)
        if ags <= 0 or (ags >= score and k != len(ops) - 1):
            split = True
            break
        mx = max(mx, ags)
    if not split and cnt["M"] >= P["minRaw"] and mx > ags:
        split = True
    counts = None
    if split:
        verdict = SPLIT
    elif cnt["M"] < P["minRaw"]:
        verdict = DROP
    else:
        tot = (cnt["M"] + cnt["R"] + cnt["I"] + cnt["D"]) & 0xFFFF
        counts = (cnt["M"] & 0xFFFF, cnt["R"] & 0xFFFF, (cnt["I"] + cnt["D"]) & 0xFFFF, tot, ags & 0xFFFF)
        verdict = SCORED if (cnt["M"] & 0xFFFF) / tot >= float(np.float32(P["minId"])) else DROP     # (-P is a float in the reference)
    frag = dict(sq=sq & 0xFFFF, eq=eq & 0xFFFF, sr=sr & 0xFFFFFFFF, rl=(er - sr + 1) & 0xFFFF)
    return [tuple(o) for o in ops], frag, score, verdict, counts


def random_runs(rng, n, tame=False):
    """n runs, no two neighbours with the same code (a DP answer comes with equal neighbours merged).  tame: long matches
    broken by short edits, the shape of a real answer (keeps the running score positive, so not every clump splits)."""
    out, prev = [], None
    for k in range(n):
        if tame:
            code = "M" if k % 2 == 0 else str(rng.choice(list("RID")))
            ln = int(rng.integers(8, 40)) if code == "M" else int(rng.integers(1, 3))
        else:
            code = str(rng.choice([c for c in "MRID" if c != prev]))
            ln = int(rng.integers(1, 40))
        out.append([code, ln])
        prev = code
    return out


def make_case(rng, P):
    max_roff = int(rng.integers(3000, 6000))
    genome = rng.integers(0, 4, size=max_roff + 64).astype(np.uint8)
    tame = rng.random() < 0.6
    tiny = rng.random() < 0.15                              # a short lonely seed: below -M unless it extends
    L = int(rng.integers(24, 60)) if tiny else int(rng.integers(60, 900))
    where = int(rng.integers(0, 5))
    r0 = [int(rng.integers(0, 3)), max_roff - L - int(rng.integers(0, 3)), int(rng.integers(300, max_roff - L - 300))][min(where, 2)]
    r0 = max(0, r0)
    q = genome[r0:r0 + L].copy()
    mut = rng.random(L) < (0.5 if tiny else rng.choice([0.0, 0.03, 0.12]))
    q[mut] = (q[mut] + rng.integers(1, 4, size=int(mut.sum()))) % 4
    # seed fragments: disjoint query intervals in order; the reference side drifts by small indels between them
    npieces = 1 if tiny else int(rng.integers(1, 13))
    cuts = np.sort(rng.choice(np.arange(2, L - 2), size=min(2 * npieces, L - 4), replace=False))
    cuts = cuts[:2 * (len(cuts) // 2)]
    if rng.random() < 0.3:
        cuts[0] = 0                                         # clump that starts at the first base of the read
    if rng.random() < 0.3:
        cuts[-1] = L - 1
    path, drift = [], 0
    for k in range(0, len(cuts), 2):
        sq, eq = int(cuts[k]), int(cuts[k + 1])
        if path:
            qgap = sq - path[-1]["eq"] - 1
            drift += int(rng.integers(-min(qgap, 3), 4)) if rng.random() < 0.6 else 0
        sr = r0 + sq + drift
        if sr < 0 or sr + (eq - sq) >= max_roff:
            break
        path.append(dict(sq=sq, eq=eq, sr=sr, rl=eq - sq + 1))
    if not path:
        return None
    answers, gaps = [], []
    for a in range(len(path) - 1):
        f1, f2 = path[a], path[a + 1]
        qgap = f2["sq"] - f1["eq"] - 1
        rgap = f2["sr"] - (f1["sr"] + f1["rl"] - 1) - 1
        if rgap < 0:
            return None
        if qgap == 0 and rgap == 0:
            continue
        g = dict(after=a, job=None, code=None, len=0, score=0)
        if qgap == 0:
            g.update(code="D", len=rgap, score=-(P["GOC"] + rgap * P["GEC"]))
        elif rgap == 0:
            g.update(code="I", len=qgap, score=-(P["GOC"] + qgap * P["GEC"]))
        elif qgap == 1 and rgap == 1:
            g.update(code="R", len=1, score=-P["RC"])
        else:
            g["job"] = len(answers)
            answers.append(dict(score=int(rng.integers(5, 40)) if tame else int(rng.integers(-60, 120)), aq=0, ar=0,
                                ops=random_runs(rng, int(rng.integers(1, 9)), tame)))
        gaps.append(g)
    # the plan of phase 1 (prepare_clumps.h): lengths left after the perfect pre-extension
    sq, sr = path[0]["sq"], path[0]["sr"]
    eq, er = path[-1]["eq"], path[-1]["sr"] + path[-1]["rl"] - 1
    back = min(sq, sr)
    m = 0
    while m < back and q[sq - 1 - m] == genome[sr - 1 - m]:
        m += 1
    back -= m
    forw = min((L - 1) - eq, max_roff - er)
    m = 0
    while m < forw and q[eq + 1 + m] == genome[er + 1 + m]:
        m += 1
    forw -= m
    prep = dict(backLen=back, forwLen=forw, jobB=None, jobF=None)
    for key, room in (("jobB", back), ("jobF", forw)):
        if room >= P["minExt"]:
            prep[key] = len(answers)
            sc = int(rng.integers(-10, 90)) if not tiny else int(rng.integers(-10, 3))
            aq = int(rng.integers(1, room + 1)) if sc > 0 else 0
            ar = max(0, aq + int(rng.integers(-2, 3))) if sc > 0 else 0
            answers.append(dict(score=sc, aq=aq, ar=ar, ops=random_runs(rng, int(rng.integers(1, 7)), tame) if sc > 0 else []))
    return genome, q, path, gaps, prep, answers, max_roff


def run_header(P, genome, q, path, gaps, prep, answers, max_roff):
    lib = S.oracle()
    lib.orc_assemble_clump.restype = C.c_int
    lib.orc_assemble_clump.argtypes = [C.c_int] * 6 + [C.c_uint32, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                        C.c_void_p]
    fr = np.zeros(len(path), S.FRAG_DT)
    for k, f in enumerate(path):
        fr[k] = (f["sr"], f["sq"], f["eq"], 0, f["rl"])
    gr = np.zeros(max(len(gaps), 1), GAP_DT)
    for k, g in enumerate(gaps):
        gr[k] = (NONE if g["job"] is None else g["job"], g["score"], g["after"], g["len"], ord(g["code"]) if g["code"] else 0, 0, 0)
    pr = np.zeros(1, PREP_DT)
    pr[0] = (0, NONE if prep["jobB"] is None else prep["jobB"], NONE if prep["jobF"] is None else prep["jobF"], len(gaps),
             prep["backLen"], prep["forwLen"], 0)
    res = np.zeros(max(len(answers), 1), RES_DT)
    flat = []
    for k, a in enumerate(answers):
        res[k] = (a["score"], a["aq"], a["ar"], len(flat) + 3, len(a["ops"]))      # (+3: answers do not start at the array's head)
        flat += a["ops"]
    rops = np.zeros(len(flat) + 4, S.OP_DT)
    for k, (code, ln) in enumerate(flat):
        rops[3 + k] = (ln, ord(code), 0)
    cap = len(path) + len(flat) + len(gaps) + 8
    out = np.zeros(cap, S.OP_DT)
    rec = np.zeros(1, ASM_DT)
    bases = pack(genome)
    qc = np.ascontiguousarray(q)
    rc = lib.orc_assemble_clump(P["GOC"], P["GEC"], P["RC"], P["MS"], P["minExt"], P["minRaw"], max_roff, float(np.float32(P["minId"])),
                                S.ptr(bases), S.ptr(qc), len(q), S.ptr(fr), len(path), S.ptr(gr), len(gaps), S.ptr(pr),
                                S.ptr(res), S.ptr(rops), S.ptr(out), cap, S.ptr(rec))
    return rc, out, rec[0]


PARAMS = [dict(GOC=5, GEC=2, RC=3, MS=1, minExt=5, minRaw=25, minId=0.9),
          dict(GOC=0, GEC=1, RC=1, MS=1, minExt=3, minRaw=15, minId=0.5),
          dict(GOC=8, GEC=3, RC=5, MS=3, minExt=3, minRaw=40, minId=0.99),
          dict(GOC=10, GEC=1, RC=4, MS=2, minExt=4, minRaw=12, minId=0.7)]


@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_header_equals_the_reference_order_of_operations(pi):
    P = PARAMS[pi]
    rng = np.random.default_rng(100 + pi)
    seen = {DROP: 0, SCORED: 0, SPLIT: 0}
    done = 0
    while done < 700:
        case = make_case(rng, P)
        if case is None:
            continue
        done += 1
        want_ops, want_frag, want_score, want_verdict, want_counts = expected(P, *case)
        rc, out, rec = run_header(P, *case)
        assert rc == 0
        got_ops = [(chr(int(o["opcode"])), int(o["length"])) for o in out[:int(rec["n_ops"])]]
        assert got_ops == want_ops, (done, case[2], case[3])
        f = rec["frag"]
        assert dict(sq=int(f["startQueryOff"]), eq=int(f["endQueryOff"]), sr=int(f["startRefOff"]), rl=int(f["refLen"])) == want_frag
        assert int(rec["score"]) == want_score
        assert int(rec["verdict"]) == want_verdict
        if want_verdict == SCORED:
            assert tuple(int(rec[k]) for k in ("matchedBases", "mismatchedBases", "gapBases", "totLength", "totScore")) == want_counts
        seen[want_verdict] += 1
    assert all(v > 0 for v in seen.values()), seen         # every outcome is exercised


def test_diverging_plan_is_reported():
    P = PARAMS[0]
    rng = np.random.default_rng(5)
    n = 0
    while n < 50:
        case = make_case(rng, P)
        if case is None:
            continue
        genome, q, path, gaps, prep, answers, max_roff = case
        if prep["jobB"] is None and prep["jobF"] is None:
            continue
        bad = dict(prep)
        if bad["jobB"] is not None:
            bad["backLen"] += 1
        else:
            bad["forwLen"] += 1
        rc, _, _ = run_header(P, genome, q, path, gaps, bad, answers, max_roff)
        assert rc == -1
        n += 1
