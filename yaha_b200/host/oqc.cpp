// oqc.cpp -- Optimal Query Coverage, filter-by-similarity and duplicate removal on the host
// (north_star: "OQC breakpoint search, FBS filtering ... stay in the host C code").
//
// Follows: clump graph nodes, key, RNG-tie-broken quicksort   GraphPath.cpp:298-459
//          deleteSubsumedDups                                  GraphPath.cpp:461-517
//          filterBySimilarity + mapping quality                GraphPath.cpp:526-692
//          accurate overlap scoring                            GraphPath.cpp:704-878
//          postFilterBySimilarity (best path search)           GraphPath.cpp:897-1086
//          postFilterRemoveDups                                GraphPath.cpp:1098-1174
//          xorshift RNG                                        Math.c:274-284
// Field widths (16-bit scores / offsets) are kept because they are observable through
// truncation in the reference.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "host.hpp"

namespace yh {

uint32_t RandState::bits()
{
    uint32_t t = s[0] ^ (s[0] >> 7);
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3]; s[3] = s[4];
    s[4] = (s[4] ^ (s[4] << 6)) ^ (t ^ (t << 13));
    return (s[1] + s[1] + 1) * s[4];
}

static const int kWorst = -(0x7fffff00);

struct CNode {                        // cGraphNode, GraphPath.cpp:299-324
    int      prev;                    // index of best predecessor in the node array, -1 none
    Clump   *clump;                   // nullptr = dead
    int16_t  bestScore, pathLength;
    uint32_t SRO, ERO;
    uint16_t SQO, EQO;                // plus-strand normalised
    int16_t  nodeLength, nodeScore;
    uint16_t qLenInOQC;
    uint8_t  reversed, seqNum;
};

static void seedRandom(ReadCtx &rc)                                     // generateRandomSeed, QueryState.c:172-187
{
    const std::vector<uint8_t> &c = rc.read->fcode;
    size_t q = 0;
    for (int i = 0; i < 5; i++) {
        uint32_t word = 0;
        for (int j = 0; j < 16; j++) { word = (word << 2) | (c[q] & 3u); if (++q >= c.size()) q = 0; }
        rc.rng.s[i] = word;
    }
}

static inline uint64_t compareKey(const CNode &n)                      // GraphPath.cpp:377-380
{
    return (((((uint64_t)n.SQO) << 16) + ((uint16_t)-(int16_t)n.EQO)) << 16) + ((uint16_t)-n.nodeScore);
}

static inline bool nodeLess(const CNode &a, const CNode &b, RandState &rng)   // :382-388
{
    uint64_t k1 = compareKey(a), k2 = compareKey(b);
    if (k1 == k2) return (rng.bits() & 1u) != 0;
    return k1 < k2;
}

static void quickSort(CNode *a, int left, int right, RandState &rng)   // GraphPath.cpp:427-453
{
    if (left >= right) return;
    int pivot = (left + right) / 2;
    std::swap(a[pivot], a[right]);
    int store = left;
    for (int i = left; i < right; i++)
        if (nodeLess(a[i], a[right], rng)) { std::swap(a[i], a[store]); store++; }
    std::swap(a[store], a[right]);
    quickSort(a, left, store - 1, rng);
    quickSort(a, store + 1, right, rng);
}

template <bool forward>
static int scoreForLength(const OpList &l, int length, const Args &A)  // GraphPath.cpp:705-732
{
    int QLen = 0, AGS = 0;
    const int n = (int)l.v.size();
    for (int k = forward ? 0 : n - 1; k >= 0 && k < n && QLen < length; k += forward ? 1 : -1) {
        const Op &o = l.v[k];
        int len = o.len;
        if (o.code == 'D') AGS -= (A.GOCost + A.GECost * len);
        else {
            if (QLen + len > length) len = length - QLen;
            QLen += len;
            if (o.code == 'M') AGS += A.MScore * len;
            else if (o.code == 'R') AGS -= A.RCost * len;
            else if (o.code == 'I') AGS -= (A.GOCost + A.GECost * len);
        }
    }
    return AGS;
}

static int accurateOverlapScore(PVec<CNode> &g, int left, int right, int overlap, const Args &A, bool *rightBest)
{                                                                       // GraphPath.cpp:744-800
    const CNode &R = g[right];
    int rightScore = R.reversed ? scoreForLength<false>(R.clump->ops, overlap, A) : scoreForLength<true>(R.clump->ops, overlap, A);
    int pathScore = 0, remaining = overlap, cur = left;
    for (;;) {
        const CNode &C = g[cur];
        int take = std::min<int>(remaining, C.qLenInOQC);
        remaining -= take;
        pathScore += C.reversed ? scoreForLength<true>(C.clump->ops, take, A) : scoreForLength<false>(C.clump->ops, take, A);
        if (remaining <= 0) break;
        cur = C.prev;
    }
    if (pathScore > rightScore) { *rightBest = false; return rightScore; }
    *rightBest = true;
    return pathScore;
}

static void cacheQlenReverse(PVec<CNode> &g, int left, int right, int overlap, bool rightBest)   // :802-826
{
    CNode &R = g[right];
    if (rightBest) {
        R.qLenInOQC = (uint16_t)(1 + R.EQO - R.SQO);
        int remaining = overlap, cur = left;
        for (;;) {
            CNode &C = g[cur];
            int take = std::min<int>(remaining, C.qLenInOQC);
            C.qLenInOQC = (uint16_t)(C.qLenInOQC - take);
            remaining -= take;
            if (remaining <= 0) break;
            cur = C.prev;
        }
    } else R.qLenInOQC = (uint16_t)((1 + R.EQO - R.SQO) - overlap);
}

static int cacheQlenPath(PVec<CNode> &g, int right, const Args &A)     // GraphPath.cpp:841-867
{
    CNode &R = g[right];
    int qLen = 1 + R.EQO - R.SQO;
    if (R.prev < 0) { R.qLenInOQC = (uint16_t)qLen; return right; }
    int left = cacheQlenPath(g, R.prev, A);
    int overlap = (int)calcOverlap(g[left].EQO, R.SQO);
    if (overlap > 0) {
        bool rb;
        accurateOverlapScore(g, left, right, overlap, A, &rb);
        cacheQlenReverse(g, left, right, overlap, rb);
    } else R.qLenInOQC = (uint16_t)qLen;
    return right;
}

struct PrimaryAttr { int alignedQueryLength, numOutputSecondaries; int16_t secondScore, thirdScore; };

static void filterBySimilarity(const Env &E, ReadCtx &rc, PVec<CNode> &g, int nodeCount, int best)
{                                                                       // GraphPath.cpp:571-692
    const Args &A = *E.A;
    std::vector<Clump *> &out = rc.scratch;
    out.clear();
    const int primeCount = g[best].pathLength;
    PVec<CNode> primaries((size_t)primeCount);
    PVec<PrimaryAttr> PA((size_t)primeCount);
    int pi = primeCount - 1;
    for (int p = best; p >= 0; p = g[p].prev) {
        primaries[pi] = g[p];
        PA[pi].alignedQueryLength = 1 + g[p].EQO - g[p].SQO;
        PA[pi].secondScore = 0; PA[pi].thirdScore = 0; PA[pi].numOutputSecondaries = 0;
        Clump *c = g[p].clump;
        c->set(kPrimary, true);
        c->matchedPrimary = (uint16_t)(pi + 1);
        out.push_back(c);
        g[p].clump = nullptr;
        pi--;
    }
    const double targetOverlap = A.FBS_PSLength;
    for (int i = 0; i < nodeCount; i++) {
        CNode &cur = g[i];
        if (!cur.clump) continue;
        Clump *c = cur.clump;
        const int curSQO = cur.SQO, curEQO = cur.EQO, curQLen = 1 + curEQO - curSQO;
        int maxOverlap = 0, maxIndex = 0;
        for (int k = 0; k < primeCount; k++) {
            int ov = 1 + std::min<int>(curEQO, primaries[k].EQO) - std::max<int>(curSQO, primaries[k].SQO);
            if (ov > maxOverlap) { maxOverlap = ov; maxIndex = k; }
        }
        if (maxOverlap > 0) {
            PrimaryAttr &P = PA[maxIndex];
            if (cur.nodeScore > P.secondScore) { P.thirdScore = P.secondScore; P.secondScore = cur.nodeScore; }
            else if (cur.nodeScore > P.thirdScore) P.thirdScore = cur.nodeScore;
            const CNode &path = primaries[maxIndex];
            if (((double)cur.nodeScore) / path.nodeScore >= A.FBS_PSScore) {
                int ov = 1 + std::min<int>(curEQO, path.EQO) - std::max<int>(curSQO, path.SQO);
                double ovD = ov;
                if (ovD / curQLen >= targetOverlap && ovD / P.alignedQueryLength >= targetOverlap) {
                    P.numOutputSecondaries += 1;
                    if (A.FBS) {
                        c->matchedPrimary = (uint16_t)(maxIndex + 1);
                        out.push_back(c);
                        continue;
                    }
                }
            }
        }
        delete c;
    }
    rc.clumps.swap(out);
    rc.primaryCount = primeCount;
    for (int k = 0; k < primeCount; k++) {                              // calcMQfromPAs, :559-569
        Clump *c = primaries[k].clump;
        const PrimaryAttr &P = PA[k];
        if (P.secondScore == 0) c->mapQuality = 250;
        else {
            double ratio = std::max(((double)c->totScore - P.secondScore), 0.0) / ((double)c->totScore);
            ratio = ratio * (1.0 + std::max(((double)c->totScore - P.thirdScore), 0.0) / c->totScore) / 2.0;
            c->mapQuality = (uint8_t)((250.0 * ratio) + 0.5);
        }
        c->numSecondaries = (uint16_t)P.numOutputSecondaries;
    }
}

void postFilterBySimilarity(const Env &E, ReadCtx &rc)                  // GraphPath.cpp:897-1086
{
    const Args &A = *E.A;
    const int nodeCount = (int)rc.clumps.size();
    if (nodeCount < 1) return;
    if (nodeCount == 1) {
        Clump *c = rc.clumps[0];
        c->set(kPrimary, true); c->mapQuality = 250; c->numSecondaries = 0; c->matchedPrimary = 1;
        rc.primaryCount = 1;
        return;
    }
    PVec<CNode> g((size_t)nodeCount);
    const int L = rc.read->len();
    int cnt = 0;
    for (int k = nodeCount - 1; k >= 0; k--) {                          // walk from the list head
        Clump *c = rc.clumps[k];
        CNode &n = g[cnt++];
        n.prev = -1; n.pathLength = 1; n.clump = c;
        n.bestScore = n.nodeScore = (int16_t)(int)c->totScore;
        n.nodeLength = (int16_t)c->totLength;
        if (c->reversed()) { n.SQO = (uint16_t)((L - 1) - c->EQO()); n.EQO = (uint16_t)((L - 1) - c->SQO()); }
        else { n.SQO = c->SQO(); n.EQO = c->EQO(); }
        n.SRO = c->SRO(); n.ERO = c->ERO();
        n.reversed = c->reversed();
        n.qLenInOQC = (uint16_t)(1 + c->EQO() - c->SQO());
        n.seqNum = (uint8_t)E.G->findSeq(n.SRO);
    }
    seedRandom(rc);
    quickSort(g.data(), 0, cnt - 1, rc.rng);

    // deleteSubsumedDups, GraphPath.cpp:488-517
    int live = 0;
    for (int i = 0; i < cnt; i++) {
        if (!g[i].clump) continue;
        const CNode cur = g[i];
        if (live != i) g[live] = g[i];
        live++;
        const int thr = cur.nodeScore / 8;
        for (int j = i + 1; j < cnt; j++) {
            CNode &nx = g[j];
            if (!nx.clump) continue;
            if (nx.EQO > cur.EQO) break;
            bool subsumed = (cur.EQO > nx.EQO && nx.nodeScore < thr);
            bool dup = (cur.SRO == nx.SRO && cur.ERO == nx.ERO && cur.reversed == nx.reversed && cur.SQO == nx.SQO && cur.EQO == nx.EQO);
            if (subsumed || dup) { delete nx.clump; nx.clump = nullptr; }
        }
    }
    cnt = live;

    int bestScore = kWorst, best = -1;
    const int minNonOverlap = A.OQCMinNonOverlap, BPCost = A.BPCost, MBPL = A.maxBPLog;
    int startj = 1;
    for (int i = 0; i < cnt; i++) {
        cacheQlenPath(g, i, A);
        CNode &Lf = g[i];
        const int leftSQO = Lf.SQO, leftEQO = Lf.EQO;
        bool foundstartj = false;
        for (int j = startj; j < cnt; j++) {
            CNode &R = g[j];
            const int rightSQO = R.SQO;
            if ((rightSQO - leftSQO) >= minNonOverlap) {
                if (!foundstartj) { startj = j; foundstartj = true; }
                const int rightEQO = R.EQO;
                if ((rightEQO - leftEQO) >= minNonOverlap) {
                    int16_t newScore = (int16_t)(Lf.bestScore + R.nodeScore);
                    if (R.bestScore > newScore) goto next_j;
                    {
                        int BPP;
                        if (Lf.seqNum == R.seqNum) {
                            uint32_t distance;
                            if (Lf.SRO > R.ERO) distance = Lf.SRO - R.ERO;
                            else if (R.SRO > Lf.ERO) distance = R.SRO - Lf.ERO;
                            else distance = 0;
                            if (distance <= 10) BPP = BPCost;
                            else {
                                double lg = log10((double)distance);
                                if (lg > MBPL) lg = (double)MBPL;
                                BPP = (int)(lg * BPCost + 0.5);
                            }
                        } else BPP = MBPL * BPCost;
                        newScore = (int16_t)(newScore - BPP);
                        if (R.bestScore > newScore) goto next_j;
                        int overlap = (int)calcOverlap(leftEQO, rightSQO);
                        bool rightBest = false;
                        if (overlap > 0) {
                            newScore = (int16_t)(newScore - accurateOverlapScore(g, i, j, overlap, A, &rightBest));
                            if (R.bestScore > newScore) goto next_j;
                        }
                        if (R.bestScore < newScore || (R.prev >= 0 && Lf.pathLength < g[R.prev].pathLength)) {
                            if (overlap > 0) {                           // cacheQlenInRightNode, :873-878
                                int qLen = 1 + R.EQO - R.SQO;
                                R.qLenInOQC = (uint16_t)(rightBest ? qLen : qLen - overlap);
                            }
                            R.bestScore = newScore; R.prev = i; R.pathLength = (int16_t)(Lf.pathLength + 1);
                        }
                    }
                }
            }
        next_j:
            if (!foundstartj) startj = cnt;
        }
        if (Lf.bestScore < bestScore) continue;
        if (Lf.bestScore > bestScore || (best >= 0 && Lf.pathLength < g[best].pathLength)) { best = i; bestScore = Lf.bestScore; }
    }
    filterBySimilarity(E, rc, g, cnt, best);
}

struct DupElem { Clump *clump; uint32_t SRO; int score; };

static int cmpDup(const void *a, const void *b)                         // GraphPath.cpp:1106-1115
{
    const DupElem *x = (const DupElem *)a, *y = (const DupElem *)b;
    if (x->SRO > y->SRO) return 1;
    if (x->SRO < y->SRO) return -1;
    return y->score - x->score;
}

void postFilterRemoveDups(const Env &E, ReadCtx &rc)                    // GraphPath.cpp:1127-1174
{
    (void)E;
    const int n = (int)rc.clumps.size();
    if (n < 2) return;
    PVec<DupElem> d((size_t)n);
    int k = 0;
    for (int i = n - 1; i >= 0; i--) { d[k].clump = rc.clumps[i]; d[k].SRO = rc.clumps[i]->SRO(); d[k].score = rc.clumps[i]->totScore; k++; }
    qsort(d.data(), (size_t)n, sizeof(DupElem), cmpDup);                // the C library's qsort, like the reference
    std::vector<Clump *> &out = rc.scratch;
    out.clear();
    for (int i = 0; i < n; i++) {
        Clump *c1 = d[i].clump;
        if (!c1) continue;
        for (int j = i + 1; j < n; j++) {
            if (d[i].SRO < d[j].SRO) break;
            Clump *c2 = d[j].clump;
            if (!c2) continue;
            if (c1->SRO() == c2->SRO() && c1->SQO() == c2->SQO() && c1->EQO() == c2->EQO() && c1->ERO() == c2->ERO() &&
                c1->reversed() == c2->reversed()) { delete c2; d[j].clump = nullptr; }
        }
        out.push_back(c1);
    }
    rc.clumps.swap(out);
}

}  // namespace yh
