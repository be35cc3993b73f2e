// graph.cpp -- fragments -> clumps on the host (kept on the CPU by north_star: "the branchy
// GraphPath ... stay in the host C code").  Input is the device's stage-2 output: the
// diag-sorted surviving fragments of one strand with their region ordinals.
//
// Follows: processFragmentsGapped region loop          QueryMatch.c:224-303
//          processFragmentRangeUsingGraph               GraphPath.cpp:272-292
//          buildBestClumpFromFragmentRange              GraphPath.cpp:161-270
//          eliminateFragments / checkStartEndCoverage   QueryMatch.c:170-215
//          addFragment / insertFragment / cleanUpClump  AlignHelpers.c:48-193
#include <algorithm>
#include <string.h>
#include "host.hpp"

namespace yh {

static const int kWorst = -(0x7fffff00);

void disposeClump(Clump *c) { delete c; }

static const uint8_t *gPrefetchBases = nullptr;                       // genome bytes (set once per run): see addFragment
static void addFragment(Clump &c, const Frag &f)                      // AlignHelpers.c:48-56
{
    // the alignment phase that follows compares bases just outside both ends of every piece (perfect extensions,
    // AlignExtFrag.cpp:30-48): ask for those genome lines now, they are cache misses in a 50 MB .. 1.5 GB array
    if (gPrefetchBases) {
        __builtin_prefetch(gPrefetchBases + ((f.startRefOff - 1u) >> 1));
        __builtin_prefetch(gPrefetchBases + ((f.startRefOff + f.refLen) >> 1));
    }
    c.matchedBases = (uint16_t)(c.matchedBases + f.refLen);
    c.path.insert(c.path.begin(), f);                                 // the reference pushes at the list head
    c.path.front().hitCount = 0;
}

static void insertFragment(Clump &c, Frag &f1)                        // AlignHelpers.c:60-90
{
    if (c.path.empty()) { addFragment(c, f1); return; }
    Frag &f2 = c.path.front();
    int maxOverlap = (int)std::max(calcOverlap(f1.endQueryOff, f2.startQueryOff), calcOverlapU(fragERO(f1), f2.startRefOff));
    if (maxOverlap > 0) {
        int l1 = fragQLen(f1), l2 = fragQLen(f2);
        bool chop1 = (l1 != l2) ? (l1 < l2) : (c.path.size() == 1);
        if (chop1) { f1.endQueryOff = (uint16_t)(f1.endQueryOff - maxOverlap); f1.refLen = (uint16_t)(f1.refLen - maxOverlap); }
        else { f2.startQueryOff = (uint16_t)(f2.startQueryOff + maxOverlap); f2.startRefOff += (uint32_t)maxOverlap;
               f2.refLen = (uint16_t)(f2.refLen - maxOverlap); }
    }
    addFragment(c, f1);
}

// The reference walks a linked list and unlinks fragments; here the fragments are an array, "unlinked" ones are
// marked and squeezed out afterwards.  Within the main loop an unlinked fragment always lies between s1 and the
// anchor, and the walk continues from the anchor, so positions that are still looked at are never marked.
static void cleanUpClump(const Args &A, Clump &c)                     // AlignHelpers.c:92-193
{
    PVec<Frag> &p = c.path;
    const int END = (int)p.size();
    bool anyGone = false;
    uint8_t goneSmall[64];
    PVec<uint8_t> goneBig;
    uint8_t *gone = goneSmall;
    if (END > 64) { goneBig.assign((size_t)END, 0); gone = goneBig.data(); } else memset(goneSmall, 0, (size_t)END);
    int s1 = 0, s2 = (END > 0) ? 1 : END, s3 = (s2 < END) ? s2 + 1 : END;
    while (s2 < END && s3 < END) {
        if (fragQLen(p[(size_t)s2]) < A.wordLen) {
            int anchor = s3;
            while (fragQLen(p[(size_t)anchor]) < A.wordLen && anchor + 1 < END) ++anchor;
            uint32_t d1 = fragDiag(p[(size_t)s1]), da = fragDiag(p[(size_t)anchor]);
            if (absDiff(d1, da) <= (uint32_t)A.maxGap) {
                for (int del = s2; del != anchor; del++) {
                    uint32_t dd = fragDiag(p[(size_t)del]);
                    bool outside = (dd < d1 && dd < da) || (dd > d1 && dd > da);
                    if (!outside || std::min(absDiff(d1, dd), absDiff(dd, da)) <= (uint32_t)A.bandWidth) { gone[del] = 1; anyGone = true; }
                }
            }
            s1 = anchor; s2 = anchor + 1;
        } else { s1 = s2; s2 = s3; }
        if (s2 < END) s3 = s2 + 1;
    }
    if (anyGone) {
        size_t w = 0;
        for (int k = 0; k < END; k++) if (!gone[k]) p[w++] = p[(size_t)k];
        p.resize(w);
    }
    // first and last fragments: only dropped when they abut their neighbour (AlignHelpers.c:154-192)
    if (p.empty()) return;
    if (fragQLen(p.front()) < A.wordLen && p.size() > 1) {
        const Frag &f1 = p[0], &f2 = p[1];
        int qGap = (int)calcGap(f1.endQueryOff, f2.startQueryOff), rGap = (int)calcGapU(fragERO(f1), f2.startRefOff);
        if ((qGap == 0 && rGap <= 2 * A.bandWidth) || (rGap == 0 && qGap <= 2 * A.bandWidth)) p.erase(p.begin());
    }
    if (fragQLen(p.back()) < A.wordLen) {
        if (p.size() == 1) return;
        const Frag &f1 = p[p.size() - 2], &f2 = p.back();
        int qGap = (int)calcGap(f1.endQueryOff, f2.startQueryOff), rGap = (int)calcGapU(fragERO(f1), f2.startRefOff);
        if ((qGap == 0 && rGap <= 2 * A.bandWidth) || (rGap == 0 && qGap <= 2 * A.bandWidth)) p.pop_back();
    }
}

struct FNode {                       // fGraphNode, GraphPath.cpp:65-79 (16-bit score fields kept)
    int      prev;                   // index of best predecessor, -1 none
    Frag    *frag;
    int16_t  bestScore, pathLength;
    uint16_t pathSQO;
    uint32_t diag;
    int16_t  nodeLength;
    uint16_t SQO, EQO;
};

// The region's consumed fragments, stamped with a generation number instead of being cleared per region.
struct Coverage {
    std::vector<uint32_t> stamp; uint32_t gen = 0;
    void reset(size_t n) { if (stamp.size() < n) stamp.resize(n, 0); if (++gen == 0) { std::fill(stamp.begin(), stamp.end(), 0); gen = 1; } }
    void mark(int i) { stamp[(size_t)i] = gen; }
    bool covered(int i) const { return stamp[(size_t)i] == gen; }
};
// Query positions covered by the clumps already cut from a region (the reference keeps one flag per position,
// QueryMatch.c:177-197): a handful of intervals, so "is [a,b] untouched" is a few comparisons, not a scan.
struct CoveredSpans {
    struct Span { int lo, hi; };
    Span sp[8]; int n = 0;
    std::vector<Span> more;                                          // beyond 8 clumps per region (rare)
    void reset() { n = 0; more.clear(); }
    void mark(int lo, int hi) { if (hi < lo) return; if (n < 8) sp[n++] = Span{lo, hi}; else more.push_back(Span{lo, hi}); }
    bool free(int a, int b) const
    {
        for (int k = 0; k < n; k++) if (sp[k].lo <= b && a <= sp[k].hi) return false;
        for (const Span &x : more) if (x.lo <= b && a <= x.hi) return false;
        return true;
    }
};

static void buildBestClump(const Args &A, Frag *frags, int lo, int hi, const Coverage &used,
                           std::vector<FNode> &nodes, Clump &clump)   // GraphPath.cpp:161-270
{
    nodes.clear();
    for (int i = lo; i <= hi; i++) {
        if (used.covered(i - lo)) continue;
        Frag &f = frags[i];
        FNode n; n.prev = -1; n.pathLength = 1; n.frag = &f; n.diag = fragDiag(f);
        n.nodeLength = (int16_t)f.refLen; n.bestScore = (int16_t)(n.nodeLength * A.MScore);
        n.SQO = f.startQueryOff; n.EQO = f.endQueryOff; n.pathSQO = n.SQO;
        nodes.push_back(n);
    }
    const int nc = (int)nodes.size();
    if (nc == 0) return;
    auto before = [](const FNode &a, const FNode &b) {
        if (a.SQO != b.SQO) return a.SQO < b.SQO;
        return a.diag < b.diag;                                      // GraphPath.cpp:148-159 (a total order: keys are distinct)
    };
    if (nc <= 24) {                                                  // the usual case: plain insertion sort
        for (int a = 1; a < nc; a++) {
            const FNode x = nodes[a];
            int b = a - 1;
            while (b >= 0 && before(x, nodes[b])) { nodes[b + 1] = nodes[b]; b--; }
            nodes[b + 1] = x;
        }
    } else std::sort(nodes.begin(), nodes.end(), before);
    int bestScore = kWorst, best = -1;
    const uint32_t maxGap = (uint32_t)A.maxGap;
    for (int i = 0; i < nc; i++) {
        FNode &L = nodes[i];
        const int lSQO = L.SQO, lEQO = L.EQO;
        const uint32_t lSRO = L.diag + (uint32_t)lSQO, lERO = L.diag + (uint32_t)L.EQO;
        for (int j = nc - 1; j > i; j--) {
            FNode &R = nodes[j];
            const int rSQO = R.SQO;
            if (rSQO == lSQO) break;
            const uint32_t diagGap = absDiff(L.diag, R.diag);
            if (diagGap > maxGap) continue;
            const uint32_t rSRO = R.diag + (uint32_t)rSQO;
            if (lSRO >= rSRO) continue;
            int desert = (int)std::min(calcGap(lEQO, rSQO), calcGapU(lERO, rSRO));
            if (desert > A.maxDesert) continue;
            int maxOverlap = (int)std::max(calcOverlap(lEQO, rSQO), calcOverlapU(lERO, rSRO));
            int newbases = R.nodeLength - maxOverlap;
            if (newbases < 1) continue;
            int gapCost = diagGap > 0 ? -(A.GOCost + (int)diagGap * A.GECost) : 0;     // calcGapCost
            int newScore = L.bestScore + newbases * A.MScore + gapCost;
            if (R.bestScore > newScore) continue;
            if (R.bestScore == newScore) {
                if (R.prev < 0) continue;
                const FNode &P = nodes[R.prev];
                int diagCompare = (int)(absDiff(L.diag, R.diag) - absDiff(P.diag, R.diag));
                if (diagCompare > 0) continue;
                if (diagCompare == 0) {
                    int gapCompare = (int)(calcGap(L.EQO, R.SQO) - calcGap(P.EQO, R.SQO));
                    if (gapCompare > 0) continue;
                    if (gapCompare == 0 && L.pathSQO <= P.pathSQO) continue;
                }
            }
            R.bestScore = (int16_t)newScore; R.prev = i; R.pathLength = (int16_t)(L.pathLength + 1); R.pathSQO = L.pathSQO;
        }
        if (L.bestScore < bestScore) continue;
        bool take = L.bestScore > bestScore;
        if (!take) {                                                  // GraphPath.cpp:88-94
            const FNode &B = nodes[best];
            take = (L.EQO != B.EQO) ? (L.EQO < B.EQO) : (L.pathSQO > B.pathSQO);
        }
        if (take) { best = i; bestScore = L.bestScore; }
    }
    for (int k = best; k >= 0; k = nodes[k].prev) insertFragment(clump, *nodes[k].frag);     // GraphPath.cpp:134-146
    if ((int)clump.matchedBases < A.minMatch) { clump.path.clear(); clump.ops.clear(); clump.matchedBases = 0; clump.status = 0; }
    else cleanUpClump(A, clump);
}

void formClumps(const Env &E, ReadCtx &rc, bool rev)
{
    const Args &A = *E.A;
    gPrefetchBases = E.G->bases;
    Frag *frags = rc.frags[rev];
    const uint32_t *reg = rc.region[rev];
    const int n = rc.nFrags[rev];
    CoveredSpans coverage;
    static thread_local Coverage used;
    static thread_local std::vector<FNode> nodes;
    Clump *spare = nullptr;                                     // an empty clump waiting for a path
    const int qSlots = rc.read->len() + 1;
    int i = 0;
    while (i < n) {
        int j = i;
        while (j + 1 < n && reg[j + 1] == reg[i]) j++;
        if (j == i) {                                                 // QueryMatch.c:281-290
            if ((int)frags[i].refLen >= A.minMatch) {
                Clump *c = spare ? spare : new Clump();
                spare = nullptr;
                addFragment(*c, frags[i]);
                c->set(kReversed, rev);                               // addClump, QueryState.c:156-161
                rc.clumps.push_back(c);
            }
        } else {                                                      // GraphPath.cpp:272-292
            coverage.reset();
            used.reset((size_t)(j - i + 1));
            for (;;) {
                Clump *c = spare ? spare : new Clump();
                spare = nullptr;
                buildBestClump(A, frags, i, j, used, nodes, *c);
                if (c->path.empty()) { *c = Clump(); spare = c; break; }
                int sqo = c->SQO(), qlen = (uint16_t)(1 + c->EQO() - c->SQO());
                coverage.mark(sqo, std::min(sqo + qlen - 1, qSlots - 1));
                // eliminateFragments, QueryMatch.c:201-215 (+ :177-197)
                const int minLeft = A.minNonOverlap - 1;
                for (int k = i; k <= j; k++) {
                    if (used.covered(k - i)) continue;
                    const int SQO = frags[k].startQueryOff, EQO = frags[k].endQueryOff;
                    bool keep = false;
                    if (EQO - SQO >= minLeft)
                        keep = coverage.free(SQO, SQO + minLeft) || coverage.free(EQO - minLeft, EQO);
                    if (!keep) used.mark(k - i);
                }
                c->set(kReversed, rev);
                rc.clumps.push_back(c);
            }
        }
        i = j + 1;
    }
    delete spare;
}

}  // namespace yh
