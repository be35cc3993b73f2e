// align.cpp -- clump alignment, scoring and splitting on the host, with every DP call turned
// into a posted device job (dpSubmit / dpWait / dpGet).
//
// Follows: alignClump, collapseSFragments           AlignHelpers.c:205-300
//          scoreClump                               AlignHelpers.c:302-366
//          splitClump / splitClumpHelper            AlignHelpers.c:374-579
//          extendFragment*ToStopPerfectly           AlignExtFrag.cpp:30-48
//          extendClumpForwardReverseTemplated       AlignExtFrag.cpp:64-156
//          makeAndAlignSFragmentToFillGap           AlignExtFrag.cpp:164-234
//          findAGS{Forward,Backward}ExtensionCarefully (post-DP trimming)   SW.cpp:553-788
//          mergeEOLToFront/Back                     SW.cpp:151-261
//
// Because a DP result is a pure function of its job tuple, all gap fills of all
// clumps of a read are posted together, then all first extensions, and only the (rare) careful
// re-extensions of split pieces are demand driven.  What follows the first DP round for a clump -- splicing the
// answers between the seed pieces, both end extensions, scoreClump's walk -- is stated in csrc/assemble_clumps.h
// (plain C99 over the device's records, the body of the device kernel of row N2) and called from here.
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include "host.hpp"
#include "../csrc/assemble_clumps.h"

namespace yh {

static_assert(sizeof(Op) == sizeof(ya_op) && offsetof(Op, len) == offsetof(ya_op, length) && offsetof(Op, code) == offsetof(ya_op, opcode),
              "assemble_clumps.h writes ya_op runs straight into a clump's op array");

uint64_t gAlignProf[4];
uint64_t gVerdictProf[8];
static const bool kAlignProf = getenv("YAHA_B200_PROF") != nullptr;
static inline uint64_t rdtsc_() { unsigned lo, hi; __asm__ volatile("rdtsc" : "=a"(lo), "=d"(hi)); return ((uint64_t)hi << 32) | lo; }


void OpVec::grow(size_t want)
{
    size_t cap = 8;                                                     // power of two: exactly one pool block
    while (cap < want || cap < (size_t)cap_ * 2) cap *= 2;
    Op *q = (Op *)TlsPool::get(cap * sizeof(Op));
    memcpy(q, p_, n_ * sizeof(Op));
    if (heap()) TlsPool::put(p_, (size_t)cap_ * sizeof(Op));
    p_ = q; cap_ = (uint32_t)cap;
}

#ifdef YH_NO_POOL
void *TlsPool::get(size_t bytes) { void *p = malloc(bytes ? bytes : 1); if (!p) { fprintf(stderr, "yaha_b200: out of memory\n"); abort(); } return p; }
void  TlsPool::put(void *p, size_t) { free(p); }
#else
namespace {
struct PoolState {
    enum { kMinShift = 5, kClasses = 12, kSlab = 1 << 20 };            // 32 B .. 64 KB
    void *freeList[kClasses] = {};
    char *slab = nullptr; size_t left = 0;
};
thread_local PoolState tPool;
inline int poolClass(size_t bytes)
{
    if (bytes <= 32) return 0;
    return 64 - __builtin_clzll((unsigned long long)(bytes - 1)) - PoolState::kMinShift;
}
}
void *TlsPool::get(size_t bytes)
{
    const int k = poolClass(bytes);
    if (k >= PoolState::kClasses) { void *p = malloc(bytes); if (!p) { fprintf(stderr, "yaha_b200: out of memory\n"); abort(); } return p; }
    PoolState &P = tPool;
    if (void *p = P.freeList[k]) { P.freeList[k] = *(void **)p; return p; }
    const size_t sz = (size_t)32 << k;
    if (P.left < sz) {
        P.slab = (char *)malloc(PoolState::kSlab);                         // (never returned: the pool lives as long as its thread)
        if (!P.slab) { fprintf(stderr, "yaha_b200: out of memory\n"); abort(); }
        P.left = PoolState::kSlab;
    }
    void *p = P.slab;
    P.slab += sz; P.left -= sz;
    return p;
}
void TlsPool::put(void *p, size_t bytes)
{
    const int k = poolClass(bytes);
    if (k >= PoolState::kClasses) { free(p); return; }
    *(void **)p = tPool.freeList[k];
    tPool.freeList[k] = p;
}
#endif

void OpVec::swap(OpVec &o) noexcept
{
    if (heap() && o.heap()) { std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(cap_, o.cap_); return; }
    OpVec *a = this, *b = &o;
    if (a->heap()) std::swap(a, b);                 // a is inline; b may be either
    Op tmp[kInline];
    const uint32_t an = a->n_;
    memcpy(tmp, a->in_, an * sizeof(Op));
    if (b->heap()) { a->p_ = b->p_; a->cap_ = b->cap_; b->p_ = b->in_; b->cap_ = kInline; }
    else memcpy(a->in_, b->in_, b->n_ * sizeof(Op));
    a->n_ = b->n_;
    memcpy(b->in_, tmp, an * sizeof(Op));
    b->n_ = an;
}

void OpList::mergeToFront(OpList &src)
{
    if (src.v.empty()) return;
    if (!v.empty() && src.v.back().code == v.front().code) {
        src.v.back().len = (uint16_t)(src.v.back().len + v.front().len);
        v.erase(v.begin());
    }
    src.v.insert(src.v.end(), v.begin(), v.end());
    v.swap(src.v);
    src.v.clear();
}

void OpList::mergeToFront(const ya_op *o, int n)
{
    if (n <= 0) return;
    int drop = 0;
    uint16_t lastLen = o[n - 1].length;
    if (!v.empty() && (char)o[n - 1].opcode == v.front().code) { lastLen = (uint16_t)(lastLen + v.front().len); drop = 1; }
    OpVec nv;
    nv.reserve((size_t)n + v.size());
    for (int k = 0; k < n; k++) nv.push_back(Op{o[k].length, (char)o[k].opcode});
    nv.back().len = lastLen;
    nv.insert(nv.end(), v.begin() + drop, v.end());
    v.swap(nv);
}

void OpList::mergeToBack(const ya_op *o, int n)
{
    if (n <= 0) return;
    int from = 0;
    if (!v.empty() && v.back().code == (char)o[0].opcode) { v.back().len = (uint16_t)(v.back().len + o[0].length); from = 1; }
    v.reserve(v.size() + (size_t)(n - from));
    for (int k = from; k < n; k++) v.push_back(Op{o[k].length, (char)o[k].opcode});
}

void OpList::mergeToBack(OpList &src)
{
    if (src.v.empty()) return;
    size_t from = 0;
    if (!v.empty() && v.back().code == src.v.front().code) {
        v.back().len = (uint16_t)(v.back().len + src.v.front().len);
        from = 1;
    }
    v.insert(v.end(), src.v.begin() + from, src.v.end());
    src.v.clear();
}

static inline int opScore(const Args &A, const Op &o)
{
    switch (o.code) {
    case 'M': return A.MScore * o.len;
    case 'R': return -(A.RCost * o.len);
    case 'I': case 'D': return -(A.GOCost + A.GECost * o.len);
    default: return 0;
    }
}

static int perfectForward(const Env &E, const uint8_t *q, Frag &f, int len)      // AlignExtFrag.cpp:30-38
{
    uint16_t qOff = (uint16_t)(f.endQueryOff + 1);
    uint32_t rOff = fragERO(f) + 1;
    int n = 0;
    while (n < len && q[qOff + n] == E.G->code(rOff + (uint32_t)n)) n++;
    if (n > 0) { f.endQueryOff = (uint16_t)(f.endQueryOff + n); f.refLen = (uint16_t)(f.refLen + n); }
    return n;
}

static int perfectBackward(const Env &E, const uint8_t *q, Frag &f, int len)     // AlignExtFrag.cpp:40-48
{
    uint16_t qOff = (uint16_t)(f.startQueryOff - 1);
    uint32_t rOff = f.startRefOff - 1;
    int n = 0;
    while (n < len && q[qOff - n] == E.G->code(rOff - (uint32_t)n)) n++;
    if (n > 0) { f.startQueryOff = (uint16_t)(f.startQueryOff - n); f.startRefOff -= (uint32_t)n; f.refLen = (uint16_t)(f.refLen + n); }
    return n;
}

// ---- phase 1 of alignClump: perfect extensions between neighbours, classify every gap (closed form or a posted DP job);
// same records as ya_prepare_clumps writes (a job number is the index of the job in this pass' list)
static void alignPrepare(const Env &E, ReadCtx &rc, Clump &c, PVec<ya_gap_rec> &gaps)
{
    const Args &A = *E.A;
    const bool rev = c.reversed();
    const uint8_t *q = rc.codes(rev);
    PVec<Frag> &p = c.path;
    const int np = (int)p.size();
    for (int k = 1; k < np; k++) {                                    // AlignHelpers.c:226-237
        Frag &l = p[(size_t)k - 1], &r = p[(size_t)k];
        int gap = (int)std::min(calcGap(l.endQueryOff, r.startQueryOff), calcGapU(fragERO(l), r.startRefOff));
        gap -= perfectBackward(E, q, r, gap);
        gap -= perfectForward(E, q, l, gap);
    }
    // (AlignHelpers.c:241-246: every piece is "nM" with score n * MScore -- written at assembly)
    for (int a = 0; a + 1 < np; a++) {                                // AlignHelpers.c:251-261 + AlignExtFrag.cpp:164-234
        const Frag &f1 = p[(size_t)a], &f2 = p[(size_t)a + 1];
        uint16_t qGap = (uint16_t)calcGap(f1.endQueryOff, f2.startQueryOff);
        uint16_t rGap = (uint16_t)calcGapU(fragERO(f1), f2.startRefOff);
        if (qGap == 0 && rGap == 0) continue;
        ya_gap_rec g; g.after = (uint16_t)a; g.job = 0xFFFFFFFFu; g.score = 0; g.len = 0; g.code = 0; g.pad = 0; g.pad2 = 0;
        const uint16_t gSQO = (uint16_t)(f1.endQueryOff + 1);
        const uint32_t gSRO = fragERO(f1) + 1;
        if (qGap == 0) { g.code = 'D'; g.len = rGap; g.score = -(A.GOCost + rGap * A.GECost); }
        else if (rGap == 0) { g.code = 'I'; g.len = qGap; g.score = -(A.GOCost + qGap * A.GECost); }
        else if (rGap == 1 && qGap == 1) { g.code = 'R'; g.len = 1; g.score = -A.RCost; }
        else {
            int lenDiff = std::abs((int)qGap - (int)rGap);
            bool banded = lenDiff + A.bandWidth * 2 + 1 < (int)rGap;
            g.job = (uint32_t)dpSubmit(rc, banded ? YA_DP_BANDED : YA_DP_FULL, rev, gSRO, rGap, gSQO, qGap).slot;
        }
        gaps.push_back(g);
    }
}

// ---- end extensions: perfect part, DP jobs, (careful) application
struct ExtState { int backLen = 0, forwLen = 0; DpFuture fb, ff; bool doB = false, doF = false; };

// perfect part of extendClumpForwardReverseTemplated (AlignExtFrag.cpp:76-107)
static void extendPerfect(const Env &E, ReadCtx &rc, Clump &c, bool goBack, bool goForw, int &score, ExtState &x)
{
    const Args &A = *E.A;
    const uint8_t *q = rc.codes(c.reversed());
    Frag &f = c.sf.front().frag;
    x.backLen = x.forwLen = 0;
    if (goBack) {
        x.backLen = (int)std::min<uint32_t>(f.startQueryOff, f.startRefOff);
        if (x.backLen > 0) {
            int m = perfectBackward(E, q, f, x.backLen);
            if (m > 0) { c.ops.v.front().len = (uint16_t)(c.ops.v.front().len + m); score += m * A.MScore; x.backLen -= m; }
        }
    }
    if (goForw) {
        uint16_t qlen = (uint16_t)((rc.read->len() - 1) - f.endQueryOff);
        uint32_t rlen = E.G->maxROff - fragERO(f);
        x.forwLen = (int)std::min<uint32_t>(qlen, rlen);
        if (x.forwLen > 0) {
            int m = perfectForward(E, q, f, x.forwLen);
            if (m > 0) { c.ops.v.back().len = (uint16_t)(c.ops.v.back().len + m); score += m * A.MScore; x.forwLen -= m; }
        }
    }
    x.doB = goBack && x.backLen >= A.minExtLength;
    x.doF = goForw && x.forwLen >= A.minExtLength;
    // Both DP jobs can be posted now: the backward extension never moves the fragment's end
    // (FragsClumps.inl:81-85), so the forward job's anchor is already final.
    if (x.doB) x.fb = dpSubmit(rc, YA_DP_EXT_BWD, c.reversed(), f.startRefOff - 1, 0, f.startQueryOff - 1, x.backLen);
    if (x.doF) x.ff = dpSubmit(rc, YA_DP_EXT_FWD, c.reversed(), fragERO(f) + 1, 0, f.endQueryOff + 1, x.forwLen);
}

// The first extensions of a clump (AlignHelpers.c:264) start from the outer ends of its first and
// last fragment; the gap fills between fragments never move those ends.  So the extension jobs can be
// posted in the SAME round as the gap fills: the perfect pre-extension is evaluated here on copies, and
// repeated for real (same outcome) after the collapse.
static void extendPlanEarly(const Env &E, ReadCtx &rc, Clump &c, ExtState &x)
{
    const Args &A = *E.A;
    const uint8_t *q = rc.codes(c.reversed());
    Frag f0 = c.path.front(), fn = c.path.back();
    x.backLen = (int)std::min<uint32_t>(f0.startQueryOff, f0.startRefOff);
    if (x.backLen > 0) x.backLen -= perfectBackward(E, q, f0, x.backLen);
    uint16_t qlen = (uint16_t)((rc.read->len() - 1) - fn.endQueryOff);
    uint32_t rlen = E.G->maxROff - fragERO(fn);
    x.forwLen = (int)std::min<uint32_t>(qlen, rlen);
    if (x.forwLen > 0) x.forwLen -= perfectForward(E, q, fn, x.forwLen);
    x.doB = x.backLen >= A.minExtLength;
    x.doF = x.forwLen >= A.minExtLength;
    if (x.doB) x.fb = dpSubmit(rc, YA_DP_EXT_BWD, c.reversed(), f0.startRefOff - 1, 0, f0.startQueryOff - 1, x.backLen);
    if (x.doF) x.ff = dpSubmit(rc, YA_DP_EXT_FWD, c.reversed(), fragERO(fn) + 1, 0, fn.endQueryOff + 1, x.forwLen);
}

// SW.cpp:671-788: trim a backward extension so the running score never reaches zero
static int carefulBackward(const Args &A, const DpAnswer &r, OpList &list, int score, int &addQ, int &addR)
{
    addQ = addR = 0;
    if (r.score <= 0) return 0;
    int QLen = 0, RLen = 0, AGS = 0, maxAGS = 0, startItem = -1;
    for (int k = 0; k < r.n; k++) {
        const Op o{r.ops[k].length, (char)r.ops[k].opcode};
        if (o.code == 'M') { QLen += o.len; RLen += o.len; }
        else if (o.code == 'R') { QLen += o.len; RLen += o.len; }
        else if (o.code == 'I') QLen += o.len;
        else if (o.code == 'D') RLen += o.len;
        AGS += opScore(A, o);
        if (AGS <= 0) { AGS = 0; maxAGS = 0; QLen = 0; RLen = 0; startItem = k; }
        if (AGS > maxAGS) maxAGS = AGS;
    }
    if (AGS <= 0 || maxAGS >= AGS + score) return 0;
    list.mergeToFront(r.ops + (startItem + 1), r.n - (startItem + 1));
    addQ = QLen; addR = RLen;
    return AGS;
}

// SW.cpp:553-669: trim a forward extension at its best point if the running score hits zero
static int carefulForward(const Args &A, const DpAnswer &r, OpList &list, int score, int &addQ, int &addR)
{
    addQ = addR = 0;
    if (r.score <= 0) return 0;
    int initAGS = r.score, keep = r.n;
    addQ = r.addedQ; addR = r.addedR;
    int QLen = 0, RLen = 0, AGS = score, maxAGS = score, maxItem = -1, maxQ = 0, maxR = 0;
    for (int k = 0; k < r.n; k++) {
        const Op o{r.ops[k].length, (char)r.ops[k].opcode};
        if (o.code == 'M' || o.code == 'R') { QLen += o.len; RLen += o.len; }
        else if (o.code == 'I') QLen += o.len;
        else if (o.code == 'D') RLen += o.len;
        AGS += opScore(A, o);
        if (AGS > maxAGS) { maxAGS = AGS; maxQ = QLen; maxR = RLen; maxItem = k; }
        else if (AGS <= 0) {
            if (maxAGS <= score) { addQ = addR = 0; return 0; }
            keep = maxItem + 1;
            addQ = maxQ; addR = maxR;
            initAGS = maxAGS - score;
            break;
        }
    }
    list.mergeToBack(r.ops, keep);
    return initAGS;
}

// DP part of extendClumpForwardReverseTemplated (AlignExtFrag.cpp:109-143)
static void extendApplyCarefully(const Env &E, ReadCtx &rc, Clump &c, ExtState &x, int score)
{
    const Args &A = *E.A;
    Frag &f = c.sf.front().frag;
    if (x.doB) {
        const DpAnswer r = dpGet(rc, x.fb);
        int aq, ar;
        const int ns = carefulBackward(A, r, c.ops, score, aq, ar);
        if (ns > 0) {
            score += ns;
            f.startQueryOff = (uint16_t)(f.startQueryOff - aq);
            f.startRefOff -= (uint32_t)ar; f.refLen = (uint16_t)(f.refLen + ar);
        }
    }
    if (x.doF) {
        const DpAnswer r = dpGet(rc, x.ff);
        int aq, ar;
        const int ns = carefulForward(A, r, c.ops, score, aq, ar);
        if (ns > 0) {
            score += ns;
            f.endQueryOff = (uint16_t)(f.endQueryOff + aq);
            f.refLen = (uint16_t)(f.refLen + ar);
        }
    }
    c.sf.front().score = score;
}

static int scoreClump(const Env &E, ReadCtx &rc, Clump *c);

// AlignHelpers.c:374-557
static int splitHelper(const Env &E, ReadCtx &rc, Clump *c, int wSQO, int wEQO)
{
    const Args &A = *E.A;
    SFrag &cs = c->sf.front();
    Frag &cf = cs.frag;
    OpList &list = cs.ops;
    list.mergeToFront(c->ops);
    uint16_t sQO = 0, eQO = 0; uint32_t sRO = 0, eRO = 0;
    int matches = 0, mism = 0, ins = 0, del = 0, AGS = 0, maxAGS = -10000, maxItem = -1, minItem = -1;
    const int n = (int)list.v.size();
    for (int k = 0; k < n; k++) {
        const Op &o = list.v[k];
        if (o.code == 'M') matches += o.len; else if (o.code == 'R') mism += o.len;
        else if (o.code == 'I') ins += o.len; else if (o.code == 'D') del += o.len;
        AGS += opScore(A, o);
        if (AGS < 0) AGS = 0;
        if (AGS > maxAGS) {
            maxAGS = AGS; maxItem = k;
            eQO = (uint16_t)(cf.startQueryOff + matches + mism + ins - 1);
            eRO = cf.startRefOff + (uint32_t)(matches + mism + del) - 1;
        }
    }
    AGS = maxAGS; matches = mism = ins = del = 0;
    int maxMatch = 0;
    for (int k = maxItem; k >= 0; k--) {
        const Op &o = list.v[k];
        if (o.code == 'M') { matches += o.len; if ((int)o.len > maxMatch) maxMatch = o.len; }
        else if (o.code == 'R') mism += o.len; else if (o.code == 'I') ins += o.len; else if (o.code == 'D') del += o.len;
        AGS -= opScore(A, o);
        if (AGS <= 0) {
            minItem = k;
            sQO = (uint16_t)(eQO - (matches + mism + ins - 1));
            sRO = eRO - (uint32_t)(matches + mism + del - 1);
            break;
        }
    }
    if (maxMatch < A.wordLen) return 0;

    int retval = 0;
    auto hasSeed = [&](const OpList &l) {                             // EditOpList2Maxmatch, SW.cpp:1215-1222
        for (const Op &o : l.v) if (o.code == 'M' && (int)o.len >= A.wordLen) return true;
        return false;
    };
    auto finishChild = [&](Clump *nc) {
        if (nc->is(kScored)) {
            nc->set(kSplit, true); nc->set(kAligned, true);
            nc->set(kReversed, c->reversed());
            rc.clumps.push_back(nc);                                  // addClump
        } else delete nc;
    };
    const Frag cur = cf;                                              // offsets before this piece is cut
    if (minItem > 0) {                                                // head piece, AlignHelpers.c:463-495
        Clump *nc = new Clump();
        nc->set(kReversed, c->reversed());
        nc->sf.emplace_back();
        SFrag &ns = nc->sf.front();
        ns.ops.v.assign(list.v.begin(), list.v.begin() + minItem);
        list.v.erase(list.v.begin(), list.v.begin() + minItem);
        maxItem -= minItem;
        if (hasSeed(ns.ops)) {
            ns.frag.hitCount = 0;
            ns.frag.startQueryOff = cur.startQueryOff; ns.frag.endQueryOff = (uint16_t)(sQO - 1);
            ns.frag.startRefOff = cur.startRefOff; fragSetERO(ns.frag, sRO - 1);
            retval += splitHelper(E, rc, nc, wSQO, wEQO);
        }
        finishChild(nc);
    }
    SFrag &cs2 = c->sf.front();                                       // (references stay valid; re-read for clarity)
    OpList &list2 = cs2.ops;
    if (maxItem != (int)list2.v.size() - 1) {                         // tail piece, AlignHelpers.c:500-531
        Clump *nc = new Clump();
        nc->set(kReversed, c->reversed());
        nc->sf.emplace_back();
        SFrag &ns = nc->sf.front();
        ns.ops.v.assign(list2.v.begin() + maxItem + 1, list2.v.end());
        list2.v.erase(list2.v.begin() + maxItem + 1, list2.v.end());
        if (hasSeed(ns.ops)) {
            ns.frag.hitCount = 0;
            ns.frag.startQueryOff = (uint16_t)(eQO + 1); ns.frag.endQueryOff = cur.endQueryOff;
            ns.frag.startRefOff = eRO + 1; fragSetERO(ns.frag, fragERO(cur));
            retval += splitHelper(E, rc, nc, wSQO, wEQO);
        }
        finishChild(nc);
    }
    Frag &f = c->sf.front().frag;
    f.startQueryOff = sQO; f.endQueryOff = eQO; f.startRefOff = sRO; fragSetERO(f, eRO);
    c->sf.front().score = maxAGS;
    c->ops.mergeToFront(c->sf.front().ops);
    // careful re-extension (AlignHelpers.c:549-551, dispatch AlignExtFrag.cpp:151-156: when neither
    // direction is requested the reference still extends forward)
    bool goBack = (sQO != wSQO), goForw = (eQO != wEQO);
    bool doBack = goBack, doForw = (goBack && goForw) || !goBack;
    ExtState x;
    int score = c->sf.front().score;
    extendPerfect(E, rc, *c, doBack, doForw, score, x);
    if (x.doB || x.doF) { uint64_t p0 = kAlignProf ? rdtsc_() : 0; dpWait(rc); if (kAlignProf) rc.parked += rdtsc_() - p0; }
    extendApplyCarefully(E, rc, *c, x, score);
    c->set(kSplit, true);
    retval += scoreClump(E, rc, c);
    return retval;
}

// the split decision of scoreClump (AlignHelpers.c:302-340), without side effects
static int willSplit(const Args &A, const Clump *c)
{
    if (c->is(kScored)) return 0;
    int AGS = 0, maxAGS = 0, matches = 0;
    const int alignedScore = c->sf.front().score;
    const int n = (int)c->ops.v.size();
    for (int k = 0; k < n; k++) {
        const Op &o = c->ops.v[k];
        if (o.code == 'M') matches += o.len;
        AGS += opScore(A, o);
        if (AGS <= 0 || (AGS >= alignedScore && k != n - 1)) return 1;
        if (AGS > maxAGS) maxAGS = AGS;
    }
    return matches >= A.minRawScore && maxAGS > AGS;
}

static int scoreClump(const Env &E, ReadCtx &rc, Clump *c)             // AlignHelpers.c:302-366
{
    const Args &A = *E.A;
    if (c->is(kScored)) return 1;
    int AGS = 0, maxAGS = 0, matches = 0, mism = 0, ins = 0, del = 0;
    const int alignedScore = c->sf.front().score;
    const int n = (int)c->ops.v.size();
    bool split = false;
    for (int k = 0; k < n; k++) {
        const Op &o = c->ops.v[k];
        if (o.code == 'M') matches += o.len; else if (o.code == 'R') mism += o.len;
        else if (o.code == 'I') ins += o.len; else if (o.code == 'D') del += o.len;
        AGS += opScore(A, o);
        if (AGS <= 0 || (AGS >= alignedScore && k != n - 1)) { split = true; break; }
        if (AGS > maxAGS) maxAGS = AGS;
    }
    if (!split && matches >= A.minRawScore && maxAGS > AGS) split = true;
    if (split) {                                                       // splitClump, AlignHelpers.c:561-579
        const Frag &f = c->sf.front().frag;
        return splitHelper(E, rc, c, f.startQueryOff, f.endQueryOff);
    }
    if (matches < A.minRawScore) return 0;
    c->matchedBases = (uint16_t)matches; c->mismatchedBases = (uint16_t)mism; c->gapBases = (uint16_t)(ins + del);
    c->totLength = (uint16_t)(matches + mism + ins + del); c->totScore = (uint16_t)AGS;
    double percent = (double)c->matchedBases / c->totLength;
    if (percent < A.minIdentity) return 0;
    c->set(kScored, true);
    return 1;
}

void postProcessClumps(const Env &E, ReadCtx &rc)                       // QueryMatch.c:306-331
{
    uint64_t q0 = rdtsc_();
    std::vector<Clump *> old;                                           // (takes the fiber's spare list: both keep their capacity)
    old.swap(rc.scratch);
    old.clear();
    old.swap(rc.clumps);
    std::reverse(old.begin(), old.end());                               // reference walks from the list head
    // phase 1: perfect extensions, gap-fill jobs AND the first extension jobs, for all clumps of the read
    struct PerClump { const ya_gap_rec *g = nullptr; ya_prep_rec prep; size_t gapLo = 0; ya_asm_rec rec; bool assembled = false; };
    PVec<PerClump> pc(old.size());
    PVec<ya_gap_rec> gaps;                                       // host-made records of every clump, [gapLo, gapLo + n_gaps) each
    bool any = false;
    bool allDev = true;                                                 // phase 1 of every clump already done on the device?
    for (size_t k = 0; k < old.size(); k++) if (!old[k]->is(kAligned) && !old[k]->prep) { allDev = false; break; }
    for (size_t k = 0; allDev && k < old.size(); k++) {
        // ya_prepare_clumps ran this phase (same source: csrc/prepare_clumps.h) and the pipeline already has the
        // answers of its jobs: the records are used where they lie, their job numbers index that first result block
        Clump &c = *old[k];
        if (c.is(kAligned)) continue;
        pc[k].prep = *c.prep;
        pc[k].g = c.gapBase + c.prep->gap_first;
    }
    if (!allDev) gaps.reserve(16 * old.size() + 8);
    for (size_t k = 0; !allDev && k < old.size(); k++) {
        if (old[k]->is(kAligned)) continue;
        pc[k].gapLo = gaps.size();
        alignPrepare(E, rc, *old[k], gaps);
        ExtState x;
        extendPlanEarly(E, rc, *old[k], x);
        ya_prep_rec &pr = pc[k].prep;
        pr.gap_first = 0; pr.n_gaps = (uint16_t)(gaps.size() - pc[k].gapLo); pr.pad = 0;
        pr.backLen = (uint16_t)x.backLen; pr.forwLen = (uint16_t)x.forwLen;
        pr.jobB = x.doB ? (uint32_t)x.fb.slot : 0xFFFFFFFFu; pr.jobF = x.doF ? (uint32_t)x.ff.slot : 0xFFFFFFFFu;
        for (size_t g = pc[k].gapLo; g < gaps.size(); g++) any |= gaps[g].job != 0xFFFFFFFFu;
        any |= x.doB || x.doF;
    }
    for (size_t k = 0; !allDev && k < old.size(); k++) pc[k].g = gaps.data() + pc[k].gapLo;
    if (kAlignProf) gAlignProf[0] += rdtsc_() - q0;
    if (any) dpWait(rc);
    q0 = rdtsc_();
    // phases 2 and 3 (csrc/assemble_clumps.h): splice the gap answers between the seed pieces, collapse, perfect-extend
    // both ends, apply both extensions, and walk the runs once for scoreClump's verdict.  The extension plan is
    // re-derived on the collapsed fragment; a difference from phase 1's (the device's, normally) is fatal -- a
    // built-in parity check of the device's plan on every read.
    ac_params AP;
    AP.GOCost = E.A->GOCost; AP.GECost = E.A->GECost; AP.RCost = E.A->RCost; AP.MScore = E.A->MScore;
    AP.minExtLength = E.A->minExtLength; AP.minRawScore = E.A->minRawScore; AP.maxROff = E.G->maxROff; AP.minIdentity = E.A->minIdentity;
    const ya_dp_result *res = nullptr; const ya_op *rops = nullptr;
    dpView(rc, res, rops);
    for (size_t k = 0; k < old.size(); k++) {
        Clump &c = *old[k];
        if (c.is(kAligned)) continue;
        const ya_prep_rec &pr = pc[k].prep;
        Op *out = c.ops.v.fill(ac_ops_bound((int)c.path.size(), pc[k].g, pr.n_gaps, &pr, res, rops));
        if (ac_assemble_clump(&AP, E.G->bases, rc.codes(c.reversed()), rc.read->len(), c.path.data(), (int)c.path.size(), pc[k].g, pr.n_gaps, &pr,
                              res, rops, reinterpret_cast<ya_op *>(out), &pc[k].rec) != 0) {
            fprintf(stderr, "yaha_b200: internal error: early extension plan diverged (planned back %d forw %d jobs %d %d; device-prepared %d)\n",
                    (int)pr.backLen, (int)pr.forwLen, (int)pr.jobB, (int)pr.jobF, c.prep != nullptr);
            abort();
        }
        c.ops.v.setSize(pc[k].rec.n_ops);
        c.sf.emplace_back();                                            // the collapsed piece
        c.sf.front().frag = pc[k].rec.frag;
        c.sf.front().score = pc[k].rec.score;
        c.path.clear();
        c.set(kAligned, true);
        pc[k].assembled = true;
        if (pc[k].rec.verdict == YA_ASM_SCORED) {                       // the MD tag will read the reference under every R and D run
            const uint8_t *g0 = E.G->bases + (pc[k].rec.frag.startRefOff >> 1), *g1 = E.G->bases + (fragERO(pc[k].rec.frag) >> 1);
            for (const uint8_t *g = g0; g <= g1; g += 64) __builtin_prefetch(g);
        }
    }
    if (kAlignProf) gAlignProf[1] += rdtsc_() - q0;
    q0 = rdtsc_();
    const uint64_t parked0 = rc.parked;
    // Clumps are scored independently; a clump that has to be split parks for its re-extensions (splitHelper).  With
    // several such clumps in a read -- repeat-rich reads have hundreds -- each is scored in a child fiber, so that
    // their re-extension rounds run side by side instead of one clump after the other.  The read's clump list is
    // put together per clump afterwards, in the order the sequential walk would have appended.
    // the verdict of an assembled clump is already known (assemble_clumps.h walked its runs): scored, dropped, or "split"
    auto verdict = [&](size_t k) { return pc[k].assembled ? (int)pc[k].rec.verdict : -1; };
    auto settle = [&](size_t k) {                                       // scoreClump for clump k, appending to the read's list
        Clump *c = old[k];
        const int v = verdict(k);
        if (v == YA_ASM_SCORED) {
            const ya_asm_rec &r = pc[k].rec;
            c->matchedBases = r.matchedBases; c->mismatchedBases = r.mismatchedBases; c->gapBases = r.gapBases;
            c->totLength = r.totLength; c->totScore = r.totScore;
            c->set(kScored, true);
        } else if (v != YA_ASM_DROP) scoreClump(E, rc, c);
        if (c->is(kScored)) rc.clumps.push_back(c);
        else delete c;
    };
    if (kAlignProf) {
        int ns = 0, nk = 0, nd = 0;
        for (size_t k = 0; k < old.size(); k++) { const int v = verdict(k); ns += v == YA_ASM_SCORED; nk += v == YA_ASM_SPLIT; nd += v == YA_ASM_DROP; }
        extern uint64_t gVerdictProf[8];
        gVerdictProf[0]++; gVerdictProf[1] += old.size() == 0; gVerdictProf[2] += (nk == 0 && old.size() > 0); gVerdictProf[3] += (nk == 0 && ns <= 1);
        gVerdictProf[4] += ns; gVerdictProf[5] += nk; gVerdictProf[6] += nd; gVerdictProf[7] += (nk == 0 && ns == 1);
    }
    int nSplit = 0;
    if (old.size() >= 2) for (size_t k = 0; k < old.size(); k++) nSplit += verdict(k) < 0 ? willSplit(*E.A, old[k]) : verdict(k) == YA_ASM_SPLIT;
    if (nSplit >= 2) {
        struct Arg { decltype(settle) *fn; } arg{&settle};
        std::vector<std::vector<Clump *>> outs(old.size());
        runAsChildren(rc, (int)old.size(), [](void *p, int k) { (*((Arg *)p)->fn)((size_t)k); }, &arg, outs.data());
        for (auto &o : outs) rc.clumps.insert(rc.clumps.end(), o.begin(), o.end());
    } else {
        for (size_t k = 0; k < old.size(); k++) settle(k);
    }
    if (kAlignProf) gAlignProf[3] += rdtsc_() - q0 - (rc.parked - parked0);
    old.clear();
    old.swap(rc.scratch);                                               // hand the buffer back to the fiber
}

}  // namespace yh
