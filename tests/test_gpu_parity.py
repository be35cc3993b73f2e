"""Parity tests proper: the CUDA path, called through the C ABI, against (a) the calls the
unmodified reference made (golden dumps) and (b) the CPU oracle on seeded synthetic inputs.
Integer work throughout => bit-exact equality is required."""
import os

import numpy as np
import pytest

import support as S
import yaha_b200
from yaha_b200 import refio, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aligner(small):
    al = yaha_b200.Aligner(small.nib, small.idx, yaha_b200.Params.defaults(word_len=11), device=0)
    al.upload_read_list(small.fwd)
    yield al
    al.close()


def _res_tuple(res, ops, i):
    o, n = int(res[i]["ops_off"]), int(res[i]["ops_n"])
    return (int(res[i]["score"]), int(res[i]["addedQLen"]), int(res[i]["addedRLen"]), S.ops_to_str(ops[o:o + n]))


def _golden_jobs(small, bw):
    jobs, want = [], []
    for rec in S.parse_dump(small.dump(bw), "D"):
        _, k, qid, st, roff, rlen, qoff, qlen, score, aq, ar, ops = rec
        jobs.append((roff, small.read_id[qid], rlen, qoff, qlen, S.KIND_OF_CHAR[k], st))
        want.append((score, aq, ar, ops))
    return np.array(jobs, dtype=yaha_b200.JOB_DT), want


@pytest.mark.parametrize("mode", ["packed", "wave", "thread"])
@pytest.mark.parametrize("bw,gap", [(5, 50), (10, 100)])
def test_dp_matches_reference_calls(small, aligner, bw, gap, mode, monkeypatch):
    """Every DP call of the reference run (all four kinds, both strands, clamped at both reference
    ends, X-drop terminated and run-to-end) reproduced: score, addedQLen, addedRLen, op runs."""
    monkeypatch.setenv("YA_DP_MODE", mode)
    aligner.set_params(yaha_b200.Params.defaults(word_len=11, bw=bw, max_gap=gap))
    jobs, want = _golden_jobs(small, bw)
    res, ops = aligner.sw_batch(jobs)
    bad = [(i, tuple(jobs[i]), _res_tuple(res, ops, i), want[i]) for i in range(len(want))
           if _res_tuple(res, ops, i) != want[i]]
    assert not bad, bad[:5]
    # cell count equals the oracle's (same definition: SW.cpp:1007-1084 bodies executed)
    p = S.default_params(word_len=11, bw=bw, max_gap=gap)
    cells = 0
    for j in jobs[:400]:
        cells += S.oracle_dp(p, small.nib.bases, small.nib.max_roff,
                             small.rev[j["read"]] if j["strand"] else small.fwd[j["read"]],
                             int(j["kind"]), int(j["rOff"]), int(j["rLen"]), int(j["qOff"]), int(j["qLen"]))[4]
    aligner.counters()
    aligner.sw_batch(jobs[:400])
    assert aligner.counters().dp_cells == cells


@pytest.mark.parametrize("sorter", ["fused", "segmented", "radix"])
def test_seed_frags_match_reference_and_oracle(small, aligner, sorter, monkeypatch):
    """Stage 1+2: per strand, surviving fragments (values and order), fragCount, totalCount and
    region ids equal the oracle; the oracle equals the reference dump (test_oracle_golden)."""
    # fused: one warp / block per strand does the whole of stage 2 in shared memory (default); segmented: the kernel chain with
    # the shared-memory sort; radix: the kernel chain with the global LSD radix sort (strands above 8192 hits)
    if sorter == "radix":
        monkeypatch.setenv("YA_SEED_RADIX", "1")
    if sorter == "segmented":
        monkeypatch.setenv("YA_SEED_CHAIN", "1")
    aligner.set_params(yaha_b200.Params.defaults(word_len=11))
    aligner.upload_read_list(small.fwd)
    strands, frags, region = aligner.seed_frags()
    p = S.default_params(word_len=11)
    ref_frags = {}
    for rec in S.parse_dump(small.dump(5), "G"):
        ref_frags[(rec[1], rec[2])] = rec[3]
    nonempty = 0
    for r, name in enumerate(small.names):
        for st in (0, 1):
            _, _, total, of, oreg, keep = S.oracle_seed_frags(p, small.idx.so, small.idx.roa, small.codes(name, st))
            s = strands[2 * r + st]
            assert int(s["total_hits"]) == total, (name, st)
            assert int(s["n_frags_all"]) == len(of), (name, st)
            if (name, st) in ref_frags:
                assert len(ref_frags[(name, st)]) == len(of)
            k = keep.astype(bool)
            a, n = int(s["first"]), int(s["n_frags"])
            assert n == int(k.sum()), (name, st)
            mine = frags[a:a + n]
            for f in ("startRefOff", "startQueryOff", "endQueryOff", "refLen"):
                assert np.array_equal(mine[f], of[k][f]), (name, st, f)
            assert np.array_equal(region[a:a + n], oreg[k]), (name, st)
            nonempty += n > 0
    assert nonempty > 500


def test_seed_frags_capacity_protocol(small, aligner):
    aligner.set_params(yaha_b200.Params.defaults(word_len=11))
    s1, f1, r1 = aligner.seed_frags(frags_cap=8)      # forces YA_E_CAPACITY then a retry
    s2, f2, r2 = aligner.seed_frags()
    assert np.array_equal(f1, f2) and np.array_equal(r1, r2) and np.array_equal(s1, s2)


def test_sw_capacity_protocol(small, aligner):
    """A too-small op buffer: YA_E_CAPACITY with every result valid, ops collected by ya_sw_fetch_ops
    without repeating the DP; the outcome equals a call with a large buffer."""
    import ctypes as C
    aligner.set_params(yaha_b200.Params.defaults(word_len=11))
    aligner.upload_read_list(small.fwd)
    jobs, want = _golden_jobs(small, 5)
    jobs = jobs[:500]
    res_big, ops_big = aligner.sw_batch(jobs, ops_cap=1 << 20)
    lib = aligner.lib
    res = np.zeros(len(jobs), dtype=yaha_b200.RES_DT)
    tiny = np.zeros(8, dtype=yaha_b200.OP_DT)
    need = C.c_size_t(0)
    before = aligner.counters().dp_cells
    rc = lib.ya_sw_batch(aligner.ctx, jobs.ctypes.data, len(jobs), res.ctypes.data, tiny.ctypes.data, len(tiny), C.byref(need))
    assert rc == yaha_b200.YA_E_CAPACITY and need.value == len(ops_big)
    assert res.tobytes() == res_big.tobytes()
    assert lib.ya_sw_fetch_ops(aligner.ctx, tiny.ctypes.data, len(tiny)) == yaha_b200.YA_E_CAPACITY
    ops = np.zeros(need.value, dtype=yaha_b200.OP_DT)
    assert lib.ya_sw_fetch_ops(aligner.ctx, ops.ctypes.data, len(ops)) == yaha_b200.YA_OK
    assert ops.tobytes() == ops_big.tobytes()
    assert aligner.counters().dp_cells > 0 and before >= 0
    res2, ops2 = aligner.sw_batch(jobs, ops_cap=4)                 # the wrapper takes the same route
    assert res2.tobytes() == res_big.tobytes() and ops2.tobytes() == ops_big.tobytes()
    r0 = np.zeros(1, dtype=yaha_b200.RES_DT)
    assert lib.ya_sw_batch(aligner.ctx, jobs.ctypes.data, 1, r0.ctypes.data, ops.ctypes.data, len(ops), C.byref(need)) == yaha_b200.YA_OK
    assert lib.ya_sw_fetch_ops(aligner.ctx, ops.ctypes.data, len(ops)) == yaha_b200.YA_E_STATE


def test_perfect_extension(small, aligner):
    rng = np.random.default_rng(5)
    jobs = []
    want = []
    lib = S.oracle()
    for _ in range(2000):
        r = int(rng.integers(0, len(small.fwd)))
        st = int(rng.integers(0, 2))
        codes = small.rev[r] if st else small.fwd[r]
        L = len(codes)
        qoff = int(rng.integers(0, L))
        name = small.names[r]
        # aim at the true locus when the name carries it, else anywhere
        roff = int(rng.integers(40, small.nib.max_roff - 40))
        fwd_dir = bool(rng.integers(0, 2))
        ln = int(rng.integers(0, (L - qoff) if fwd_dir else qoff + 1))
        ln = min(ln, 30)
        jobs.append((roff, r, 0, qoff, ln, yaha_b200.DP_EXT_FWD if fwd_dir else yaha_b200.DP_EXT_BWD, st))
        want.append(lib.orc_perfect(S.ptr(small.nib.bases), S.ptr(codes), roff, qoff, ln, 1 if fwd_dir else -1))
    got = aligner.perfect_ext(np.array(jobs, dtype=yaha_b200.JOB_DT))
    assert list(got) == want


def test_random_jobs_against_oracle(small, aligner):
    """Seeded random job tuples (not only those the pipeline produces): odd sizes, tiny and wide
    global jobs, extensions started at arbitrary anchors near both reference ends, non-default
    scoring -- including caps (maxGap) that bind inside the band."""
    rng = np.random.default_rng(11)
    for (bw, gap, goc, gec, rc, ms, x) in [(5, 50, 5, 2, 3, 1, 25), (3, 4, 2, 1, 2, 1, 10), (8, 30, 6, 1, 4, 2, 40),
                                            (1, 2, 1, 1, 1, 1, 5), (16, 100, 5, 2, 3, 1, 25)]:
        P = yaha_b200.Params.defaults(word_len=11, bw=bw, max_gap=gap, goc=goc, gec=gec, rc=rc, ms=ms, x=x)
        aligner.set_params(P)
        p = S.default_params(word_len=11, bw=bw, max_gap=gap, goc=goc, gec=gec, rc=rc, ms=ms, x=x)
        jobs = []
        for _ in range(300):
            r = int(rng.integers(0, 450))
            st = int(rng.integers(0, 2))
            L = len(small.fwd[r])
            kind = int(rng.integers(0, 4))
            # true locus from the read name r<i>_<start>_<strand>
            parts = small.names[r].split("_")
            start = int(parts[1])
            gstart = start + (0 if start < 120000 else (0 if start < 200003 else 5)) + (0 if start < 200003 else 0)
            if kind <= 1:
                qlen = int(rng.integers(1, 60)); rlen = max(1, qlen + int(rng.integers(-min(gap, qlen - 1) if qlen > 1 else 0, gap + 1)))
                qoff = int(rng.integers(0, L - qlen))
                roff = max(0, min(small.nib.max_roff - rlen - 1, gstart + qoff + int(rng.integers(-3, 4))))
                jobs.append((roff, r, rlen, qoff, qlen, kind, st))
            else:
                where = rng.integers(0, 4)
                if kind == yaha_b200.DP_EXT_FWD:
                    qoff = int(rng.integers(0, L)); qlen = L - qoff
                    roff = [gstart + qoff, small.nib.max_roff - int(rng.integers(1, 60)), gstart + qoff + 2, int(rng.integers(0, 50))][where]
                else:
                    qoff = int(rng.integers(0, L)); qlen = qoff + 1
                    roff = [gstart + qoff, int(rng.integers(0, 60)), gstart + qoff - 2, small.nib.max_roff - int(rng.integers(1, 50))][where]
                roff = max(0, min(small.nib.max_roff - 1, roff))
                jobs.append((roff, r, 0, qoff, qlen, kind, st))
        jobs = np.array(jobs, dtype=yaha_b200.JOB_DT)
        res, ops = aligner.sw_batch(jobs)
        for i, j in enumerate(jobs):
            codes = small.rev[j["read"]] if j["strand"] else small.fwd[j["read"]]
            w = S.oracle_dp(p, small.nib.bases, small.nib.max_roff, codes, int(j["kind"]), int(j["rOff"]),
                            int(j["rLen"]), int(j["qOff"]), int(j["qLen"]))[:4]
            assert _res_tuple(res, ops, i) == w, (tuple(j), (bw, gap, goc, gec, rc, ms, x))


def test_packed_kernel_score_range_guard(small):
    """dp_ext_packed_kernel keeps scores x256 in int32 under a -2^29 sentinel; ya_sw_batch must route a job whose
    (rows + W + 2) * largest step cost reaches 2^20 to dp_wave_kernel (plain int32 like SW.cpp).  Long extensions with
    costs at the top of what the reference's own arithmetic allows (GOC + GEC <= 256, SW.cpp:356) against the oracle;
    jobs below the bound in the same batch stay on the packed kernel."""
    b = small.nib.bases
    n = 9000
    packed = np.asarray(b[500:500 + n // 2 + 1])
    codes = np.empty(2 * len(packed), np.uint8)
    codes[0::2] = packed >> 4; codes[1::2] = packed & 15
    exact = np.ascontiguousarray(codes[:n])                         # reference bases 1000 .. 1000 + n
    rng = np.random.default_rng(5)
    noisy = exact.copy()
    hit = rng.random(n) < 0.03
    noisy[hit] = (noisy[hit] + rng.integers(1, 4, size=int(hit.sum()))) % 4
    reads = [exact, noisy, small.fwd[0]]
    for (bw, gap, goc, gec, rc, ms, x) in [(5, 50, 200, 50, 250, 30, 600), (10, 100, 150, 100, 255, 100, 2000), (5, 50, 5, 2, 3, 1, 25)]:
        P = yaha_b200.Params.defaults(word_len=11, bw=bw, max_gap=gap, goc=goc, gec=gec, rc=rc, ms=ms, x=x)
        al = yaha_b200.Aligner(small.nib, small.idx, P, device=0)
        al.upload_read_list(reads)
        p = S.default_params(word_len=11, bw=bw, max_gap=gap, goc=goc, gec=gec, rc=rc, ms=ms, x=x)
        jobs = []
        for r in (0, 1):
            for q0 in (0, 17, 4000, 8000):
                jobs.append((1000 + q0, r, 0, q0, n - q0, yaha_b200.DP_EXT_FWD, 0))
                jobs.append((1000 + n - 1 - q0, r, 0, n - 1 - q0, n - q0, yaha_b200.DP_EXT_BWD, 0))
            jobs.append((1000 + 8900, r, 0, 8900, 100, yaha_b200.DP_EXT_FWD, 0))      # short: below the bound
        jobs = np.array(jobs, dtype=yaha_b200.JOB_DT)
        res, ops = al.sw_batch(jobs)
        long_rows = 0
        for i, j in enumerate(jobs):
            w = S.oracle_dp(p, small.nib.bases, small.nib.max_roff, reads[int(j["read"])], int(j["kind"]), int(j["rOff"]),
                            int(j["rLen"]), int(j["qOff"]), int(j["qLen"]))[:4]
            assert _res_tuple(res, ops, i) == w, (tuple(j), (bw, gap, goc, gec, rc, ms, x))
            long_rows = max(long_rows, int(res[i]["addedQLen"]))
        assert long_rows > 4000                                    # the exact copies really run thousands of rows
        al.close()


def test_empty_and_degenerate_inputs(small):
    al = yaha_b200.Aligner(small.nib, small.idx, yaha_b200.Params.defaults(word_len=11), device=0)
    # empty batch
    al.upload_read_list([])
    strands, frags, region = al.seed_frags()
    assert len(strands) == 0 and len(frags) == 0
    # reads shorter than K, all-N reads, one real read
    reads = [np.zeros(5, np.uint8), np.full(50, 4, np.uint8), small.fwd[0], np.zeros(0, np.uint8)]
    al.upload_read_list(reads)
    strands, frags, region = al.seed_frags()
    assert list(strands["n_frags_all"][:4]) == [0, 0, 0, 0]
    assert strands["n_frags_all"][4] + strands["n_frags_all"][5] > 0
    assert list(strands["n_frags_all"][6:]) == [0, 0]
    res, ops = al.sw_batch(np.zeros(0, dtype=yaha_b200.JOB_DT))
    assert len(res) == 0
    # extension with nothing left to extend returns 0 / no ops (SW.cpp:494)
    j = np.array([(1000, 2, 0, 0, 0, yaha_b200.DP_EXT_FWD, 0)], dtype=yaha_b200.JOB_DT)
    res, ops = al.sw_batch(j)
    assert int(res[0]["score"]) == 0 and int(res[0]["ops_n"]) == 0
    with pytest.raises(yaha_b200.YahaError):
        al.sw_batch(np.array([(0, 99, 5, 0, 5, 0, 0)], dtype=yaha_b200.JOB_DT))    # bad read index
    al.close()


def test_linearity_of_sharding(small, aligner):
    """Size-independent property: results for a batch equal the concatenation of results for its
    shards (reads are independent units)."""
    aligner.set_params(yaha_b200.Params.defaults(word_len=11))
    aligner.upload_read_list(small.fwd)
    s_all, f_all, r_all = aligner.seed_frags()
    half = len(small.fwd) // 2
    aligner.upload_read_list(small.fwd[:half])
    s_a, f_a, r_a = aligner.seed_frags()
    aligner.upload_read_list(small.fwd[half:])
    s_b, f_b, r_b = aligner.seed_frags()
    assert np.array_equal(np.concatenate([f_a, f_b]), f_all)
    assert np.array_equal(np.concatenate([r_a, r_b]), r_all)
    assert np.array_equal(np.concatenate([s_a["n_frags"], s_b["n_frags"]]), s_all["n_frags"])
    aligner.upload_read_list(small.fwd)


def test_device_index_build_is_bit_identical(small):
    """ya_open_build (Index.c:49-331 replacement) must reproduce the reference's SO and ROA
    arrays exactly (the golden index digest is checked in test_formats)."""
    al = yaha_b200.Aligner(small.nib, None, yaha_b200.Params.defaults(word_len=11), device=0)
    idx = al.download_index()
    assert np.array_equal(idx.so, np.asarray(small.idx.so))
    assert np.array_equal(idx.roa, np.asarray(small.idx.roa))
    # and it aligns identically
    al.upload_read_list(small.fwd[:50])
    s1, f1, r1 = al.seed_frags()
    al.close()
    al2 = yaha_b200.Aligner(small.nib, small.idx, yaha_b200.Params.defaults(word_len=11), device=0)
    al2.upload_read_list(small.fwd[:50])
    s2, f2, r2 = al2.seed_frags()
    al2.close()
    assert np.array_equal(f1, f2) and np.array_equal(s1, s2)


def test_device_index_build_every_word_length_skip_and_hit_cap(small, tmp_path):
    """Row N3: ya_open_build for every -L / -S / -H the reference was run with (golden/small/index_variants.json:
    sha256 of the files the UNMODIFIED reference wrote for the golden reference and for an N-riddled one), i.e. the
    -S grid with its renormalisation behind runs of non-ACGT codes (Index.c:105-128) and the down-sampling of
    over-full k-mers with the reference's generator (Index.c:271-315, Math.c:274-343)."""
    import hashlib
    import json
    want = json.load(open(os.path.join(small.golden, "index_variants.json")))
    nrich = refio.build_nib2(synth.n_rich_reference())
    open(tmp_path / "nrich.nib2", "wb").write(nrich)
    nibs = {"nrich": refio.load_nib2(str(tmp_path / "nrich.nib2")), "small": small.nib}
    checked = sampled = 0
    for name, digest in want.items():
        if name.endswith(".nib2"):
            continue
        stem, spec = name.split(".X")
        L, Sk, H = int(spec[0:2]), int(spec[3:5]), int(spec[6:11])
        al = yaha_b200.Aligner(nibs[stem], None, yaha_b200.Params.defaults(word_len=L, max_hits=min(650, H)), device=0,
                               build_max_hits=H, build_skip=Sk)
        idx = al.download_index(max_hits=H)
        al.close()
        path = str(tmp_path / name)
        refio.write_index(path, idx)
        if hashlib.sha256(open(path, "rb").read()).hexdigest() != digest:      # say where (the host builder is pinned to the same digests)
            img = np.frombuffer(refio.build_index(nibs[stem], L, max_hits=H, skip=Sk), dtype="<u4")
            n_so = 4 ** L + 1
            wso, wroa = img[4:4 + n_so], img[4 + n_so:]
            bad_so = np.nonzero(idx.so != wso)[0][:5] if len(idx.so) == len(wso) else "length"
            bad_roa = np.nonzero(idx.roa != wroa)[0][:5] if len(idx.roa) == len(wroa) else ("length", len(idx.roa), len(wroa))
            raise AssertionError((name, "so differs at", bad_so, "roa differs at", bad_roa))
        sampled += int(np.max(np.diff(idx.so.astype(np.int64))) == H and H < 65525)
        checked += 1
    assert checked == 14 and sampled >= 2


_AB_SNIPPET = r"""
import sys, os
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import yaha_b200, support as S
from conftest import Small
small = Small({tmp!r})
out = []
for bw, gap in ((5, 50), (10, 100)):
    al = yaha_b200.Aligner(small.nib, small.idx, yaha_b200.Params.defaults(word_len=11, bw=bw, max_gap=gap), device=0)
    al.upload_read_list(small.fwd)
    jobs = []
    for rec in S.parse_dump(small.dump(bw), "D"):
        _, k, qid, st, roff, rlen, qoff, qlen, score, aq, ar, ops = rec
        jobs.append((roff, small.read_id[qid], rlen, qoff, qlen, S.KIND_OF_CHAR[k], st))
    res, ops = al.sw_batch(np.array(jobs, dtype=yaha_b200.JOB_DT))
    ops = np.asarray(ops)
    for i in range(len(res)):            # (placement of a job's runs in ops[] is arbitrary: serialise per job)
        o, n = int(res[i]["ops_off"]), int(res[i]["ops_n"])
        out.append(np.array([res[i]["score"], res[i]["addedQLen"], res[i]["addedRLen"], n], dtype=np.int64).tobytes())
        out.append(ops[o:o + n].tobytes())
    al.close()
open({dst!r}, "wb").write(b"".join(out))
"""


@pytest.mark.gpu
def test_kernel_variants_agree_bytewise(small, tmp_path):
    """The environment switches select kernels per process, so each variant runs in its own interpreter on the
    reference's DP calls: warp-per-job vs thread-per-job traceback, full-matrix gap fills on the wavefront kernel
    vs one thread per job.  Results (scores, lengths, op arrays) must be byte-identical."""
    import subprocess
    import sys
    blobs = {}
    # (the packed X-drop kernel has three geometries chosen by launch size -- 8/16, 4/8 and 2/4 jobs per warp: the golden calls
    #  are a small launch, i.e. the narrowest by default; the two thresholds force the other two)
    for name, env in (("default", {}), ("tb_thread", {"YA_TB": "thread"}), ("full_thread", {"YA_FULL_THREAD_MAXW": "100000"}),
                      ("packed_wide", {"YA_PACKED_NARROW_BELOW": "0", "YA_PACKED_XNARROW_BELOW": "0"}),
                      ("packed_narrow", {"YA_PACKED_NARROW_BELOW": "100000000", "YA_PACKED_XNARROW_BELOW": "0"}),
                      # reference windows staged into shared memory by cp.async.bulk (TMA) instead of read through L1/L2
                      ("staged", {"YA_EXT_STAGE": "1"}),
                      ("staged_wide", {"YA_EXT_STAGE": "1", "YA_PACKED_NARROW_BELOW": "0", "YA_PACKED_XNARROW_BELOW": "0"})):
        dst = str(tmp_path / f"{name}.bin")
        code = _AB_SNIPPET.format(root=S.ROOT, tmp=str(tmp_path / f"small_{name}"), dst=dst)
        os.makedirs(str(tmp_path / f"small_{name}"), exist_ok=True)
        p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        blobs[name] = open(dst, "rb").read()
    assert blobs["default"] == blobs["tb_thread"]
    assert blobs["default"] == blobs["full_thread"]
    assert blobs["default"] == blobs["packed_wide"]
    assert blobs["default"] == blobs["packed_narrow"]
    assert blobs["default"] == blobs["staged"]
    assert blobs["default"] == blobs["staged_wide"]
    assert len(blobs["default"]) > 100000


def test_device_clumps_match_cpu_build_of_the_same_source(small, aligner):
    """Row N1: ya_form_clumps (one thread per strand running csrc/form_clumps.h) against the CPU build of the same
    source (oracle/oracle_clumps.c) on the fragments the device itself produced -- clump by clump, fragment by
    fragment.  The source is pinned through the host program (it is its formClumps; golden SAMs)."""
    import ctypes as C
    lib = S.oracle()
    for bw, gap, desert, mno in ((5, 50, 50, 25), (10, 100, 50, 25), (5, 50, 10, 40)):
        aligner.set_params(yaha_b200.Params.defaults(word_len=11, bw=bw, max_gap=gap))
        strands, frags, region = aligner.seed_frags()
        first, count, clumps, path = aligner.form_clumps(len(frags), max_desert=desert, min_non_overlap=mno)
        n_checked = 0
        for s in range(len(strands)):
            n, f0 = int(strands[s]["n_frags"]), int(strands[s]["first"])
            if n == 0:
                assert count[s] == 0
                continue
            assert count[s] != 0xFFFFFFFF
            L = len(small.fwd[s >> 1])
            fr = np.ascontiguousarray(frags[f0:f0 + n]); rg = np.ascontiguousarray(region[f0:f0 + n])
            opath = np.zeros(n, dtype=yaha_b200.FRAG_DT); ocl = np.zeros(n, dtype=yaha_b200.CLUMP_DT)
            nc = lib.orc_form_clumps(11, gap, desert, 25, mno, bw, 5, 2, 1, S.ptr(fr), S.ptr(rg), n, L, S.ptr(opath), S.ptr(ocl))
            assert nc == int(count[s]), (s, nc, int(count[s]))
            c0 = int(first[s])
            for k in range(nc):
                mine, want = clumps[c0 + k], ocl[k]
                assert int(mine["n"]) == int(want["n"]) and int(mine["matchedBases"]) == int(want["matchedBases"]), (s, k)
                a = path[int(mine["first"]):int(mine["first"]) + int(mine["n"])]
                b = opath[int(want["first"]):int(want["first"]) + int(want["n"])]
                assert a.tobytes() == b.tobytes(), (s, k)
                n_checked += 1
        assert n_checked > 300
    aligner.set_params(yaha_b200.Params.defaults(word_len=11))


def test_align_batch_through_the_abi_gives_the_reference_sam(small):
    """Rows N2 / N4 through the C ABI (ctypes, no host program): ya_align_batch takes the golden reads as text and must return,
    for every read it does not hand back, exactly the lines the UNMODIFIED reference wrote for that read (golden SAMs, default
    flags and -FBS Y), in the reference's order; the reads it hands back are few."""
    import gzip
    for golden, reads_file, fbs in (("out_bw5.sam.gz", "reads.fa", False), ("out_multi.sam.gz", "multi.fa", False), ("out_fbs.sam.gz", "reads.fa", True)):
        want = {}
        for line in gzip.open(os.path.join(small.golden, golden), "rt"):
            if not line.startswith("@"):
                want.setdefault(line.split("\t", 1)[0], []).append(line)
        reads = [(n, bytes(s)) for n, s in refio.read_queries(os.path.join(small.dir, reads_file), word_len=11)]
        al = yaha_b200.Aligner(small.nib, small.idx, yaha_b200.Params.defaults(word_len=11), device=0)
        al.set_output(yaha_b200.OutParams.defaults(fbs=fbs))
        text, toff, status = al.align_batch(reads)
        c = al.counters()
        al.close()
        assert int(status.sum()) <= len(reads) // 20 and c.reads_finished == len(reads) - int(status.sum())
        n_lines = 0
        for r, (name, _) in enumerate(reads):
            mine = text[int(toff[r]):int(toff[r + 1])].decode()
            if status[r]:
                assert mine == ""
                continue
            assert mine == "".join(want.get(name, [])), name
            n_lines += len(want.get(name, []))
        assert n_lines > 400


def test_align_batch_capacity_protocol_and_degenerate_batches(small):
    """ya_align_batch: a text buffer that is too small returns YA_E_CAPACITY with text_needed, offsets and status valid and the
    text fetched afterwards (ya_align_fetch_text) identical to a call with room; an empty batch and a batch of reads without a
    single hit are fine; calling before ya_set_output is a state error."""
    import ctypes as C
    reads = [(n, bytes(s)) for n, s in refio.read_queries(os.path.join(small.dir, "reads.fa"), word_len=11)][:100]
    al = yaha_b200.Aligner(small.nib, small.idx, yaha_b200.Params.defaults(word_len=11), device=0)
    with pytest.raises(yaha_b200.YahaError):
        al.align_batch(reads)
    al.set_output(yaha_b200.OutParams.defaults())
    text, toff, status = al.align_batch(reads)
    assert len(text) > 10000 and int(toff[-1]) == len(text)
    # too small a buffer: go through the ABI by hand
    n = len(reads)
    offs = np.zeros(n + 1, dtype=np.uint64); offs[1:] = np.cumsum([len(s) for _, s in reads])
    idoff = np.zeros(n + 1, dtype=np.uint32); idoff[1:] = np.cumsum([len(i) for i, _ in reads])
    chars = np.frombuffer(b"".join(s for _, s in reads), dtype=np.uint8).copy()
    ids = np.frombuffer("".join(i for i, _ in reads).encode(), dtype=np.uint8).copy()
    small_text = np.zeros(64, dtype=np.uint8); toff2 = np.zeros(n + 1, dtype=np.uint64); st2 = np.zeros(n, dtype=np.uint8)
    tb = yaha_b200._TextBatch(n, chars.ctypes.data, offs.ctypes.data, None, ids.ctypes.data, idoff.ctypes.data, small_text.ctypes.data, 64,
                              toff2.ctypes.data, st2.ctypes.data, 0, 0, 0)
    rc = al.lib.ya_align_batch(al.ctx, C.byref(tb))
    assert rc == yaha_b200.YA_E_CAPACITY and tb.text_needed == len(text)
    assert np.array_equal(toff2, toff) and np.array_equal(st2, status)
    full = np.zeros(int(tb.text_needed), dtype=np.uint8)
    assert al.lib.ya_align_fetch_text(al.ctx, full.ctypes.data, len(full)) == 0
    assert full.tobytes() == text
    # degenerate batches
    t0, o0, s0 = al.align_batch([])
    assert t0 == b"" and len(s0) == 0
    junk = [("junk%d" % k, b"ACGT" * 40) for k in range(3)] + [("nnn", b"N" * 100)]
    t1, o1, s1 = al.align_batch(junk)
    assert int(s1.sum()) == 0 and int(o1[-1]) == len(t1)
    al.close()
