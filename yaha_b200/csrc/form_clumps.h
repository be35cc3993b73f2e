/* form_clumps.h -- fragments of one strand -> clumps of seed fragments.
 *
 * ONE statement of the algorithm, plain C99, compiled three ways: as device code by clumps.cu (one warp per
 * strand, ya_form_clumps), as host code by the host program (yaha_b200/host/graph.cpp) and by the oracle-backed
 * mock of the ABI (tests/mock/mock_abi.c).  The host program's golden tests therefore pin the very code the
 * kernel runs.
 *
 * Follows: processFragmentsGapped region loop            QueryMatch.c:224-303
 *          processFragmentRangeUsingGraph                 GraphPath.cpp:272-292
 *          buildBestClumpFromFragmentRange                GraphPath.cpp:161-270  (16-bit node fields kept, :65-79)
 *          eliminateFragments / checkStartEndCoverage     QueryMatch.c:170-215
 *          addFragment / insertFragment / cleanUpClump    AlignHelpers.c:48-193
 *
 * Input : the surviving fragments of one strand in diagonal order with their region ordinals (stage 2 output).
 * Output: clumps in the reference's creation order, each a run of fragments in query order (after the overlap
 *         chops of insertFragment and the clean-up), written to out_path; fc_clump records index into it.
 * Scratch (caller provided, n = number of fragments of the strand): nodes[n], used[n], tmp[n] fragments.
 */
#ifndef YAHA_B200_FORM_CLUMPS_H
#define YAHA_B200_FORM_CLUMPS_H
#include <stdint.h>
#include "../../include/yaha_b200.h"

#ifdef __CUDACC__
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD static inline
#endif

/* On the device (clumps.cu defines FC_WARP_COOP) one WARP runs a strand: all 32 lanes execute fc_form_clumps with the same
 * arguments; the serial parts are done by lane 0 and their results broadcast, and the quadratic part -- the inner loop of
 * the chain DP, whose iterations update different nodes -- is strided over the lanes.  Everywhere else the macros below
 * make the same source plain serial C. */
#if defined(__CUDA_ARCH__) && defined(FC_WARP_COOP)
#define FC_LANE      ((int)(threadIdx.x & 31u))
#define FC_NLANES    32
#define FC_SYNC()    __syncwarp()
#define FC_BCAST(x)  ((x) = __shfl_sync(0xffffffffu, (x), 0))
#else
#define FC_LANE      0
#define FC_NLANES    1
#define FC_SYNC()    ((void)0)
#define FC_BCAST(x)  ((void)0)
#endif

typedef struct fc_params {
    int32_t wordLen, maxGap, maxDesert, minMatch, minNonOverlap, bandWidth, GOCost, GECost, MScore;
} fc_params;

typedef struct fc_node {              /* fGraphNode, GraphPath.cpp:65-79 */
    int32_t  prev;                    /* index of best predecessor, -1 none */
    int32_t  frag;                    /* index into the strand's fragment array */
    int16_t  bestScore, pathLength;
    uint16_t pathSQO;
    uint16_t SQO, EQO;
    int16_t  nodeLength;
    uint32_t diag;
} fc_node;

FC_HD int      fc_qlen(const ya_frag *f) { return 1 + (int)f->endQueryOff - (int)f->startQueryOff; }
FC_HD uint32_t fc_ero(const ya_frag *f) { return f->startRefOff + f->refLen - 1; }
FC_HD uint32_t fc_diag(const ya_frag *f) { return f->startRefOff - f->startQueryOff; }
FC_HD uint32_t fc_absdiff(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }                  /* FragsClumps.inl:133-137 */
FC_HD unsigned fc_gap(int lo, int hi) { return hi > lo ? (unsigned)(hi - lo) - 1 : 0; }               /* :157 */
FC_HD unsigned fc_overlap(int lo, int hi) { return lo >= hi ? (unsigned)(lo - hi) + 1 : 0; }          /* :158 */
FC_HD unsigned fc_gap_u(uint32_t lo, uint32_t hi) { return hi > lo ? (hi - lo) - 1 : 0; }
FC_HD unsigned fc_overlap_u(uint32_t lo, uint32_t hi) { return lo >= hi ? (lo - hi) + 1 : 0; }
FC_HD unsigned fc_min_u(unsigned a, unsigned b) { return a < b ? a : b; }
FC_HD unsigned fc_max_u(unsigned a, unsigned b) { return a > b ? a : b; }

FC_HD int fc_before(const fc_node *a, const fc_node *b)              /* GraphPath.cpp:148-159 (keys are distinct) */
{
    if (a->SQO != b->SQO) return a->SQO < b->SQO;
    return a->diag < b->diag;
}

/* Sorts nodes[0..nc) by (SQO, diag): insertion sort for the usual handful, heap sort beyond (same order: distinct keys). */
FC_HD void fc_sort_nodes(fc_node *nodes, int nc)
{
    if (nc <= 32) {
        for (int a = 1; a < nc; a++) {
            const fc_node x = nodes[a];
            int b = a - 1;
            while (b >= 0 && fc_before(&x, &nodes[b])) { nodes[b + 1] = nodes[b]; b--; }
            nodes[b + 1] = x;
        }
        return;
    }
    for (int start = nc / 2 - 1; start >= 0; start--) {              /* heapify */
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= nc) break;
            if (child + 1 < nc && fc_before(&nodes[child], &nodes[child + 1])) child++;
            if (!fc_before(&nodes[root], &nodes[child])) break;
            const fc_node t = nodes[root]; nodes[root] = nodes[child]; nodes[child] = t;
            root = child;
        }
    }
    for (int end = nc - 1; end > 0; end--) {
        const fc_node t = nodes[0]; nodes[0] = nodes[end]; nodes[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && fc_before(&nodes[child], &nodes[child + 1])) child++;
            if (!fc_before(&nodes[root], &nodes[child])) break;
            const fc_node u = nodes[root]; nodes[root] = nodes[child]; nodes[child] = u;
            root = child;
        }
    }
}

/* cleanUpClump (AlignHelpers.c:92-193) on an array: "unlinked" fragments are marked in gone[] and squeezed out.
 * Within the main loop an unlinked fragment always lies between s1 and the anchor and the walk continues from the
 * anchor, so positions that are still looked at are never marked.  Returns the new length. */
FC_HD int fc_clean_up(const fc_params *P, ya_frag *p, int END, uint8_t *gone)
{
    int anyGone = 0;
    for (int k = 0; k < END; k++) gone[k] = 0;
    int s1 = 0, s2 = (END > 0) ? 1 : END, s3 = (s2 < END) ? s2 + 1 : END;
    while (s2 < END && s3 < END) {
        if (fc_qlen(&p[s2]) < P->wordLen) {
            int anchor = s3;
            while (fc_qlen(&p[anchor]) < P->wordLen && anchor + 1 < END) ++anchor;
            const uint32_t d1 = fc_diag(&p[s1]), da = fc_diag(&p[anchor]);
            if (fc_absdiff(d1, da) <= (uint32_t)P->maxGap) {
                for (int del = s2; del != anchor; del++) {
                    const uint32_t dd = fc_diag(&p[del]);
                    const int outside = (dd < d1 && dd < da) || (dd > d1 && dd > da);
                    if (!outside || fc_min_u(fc_absdiff(d1, dd), fc_absdiff(dd, da)) <= (uint32_t)P->bandWidth) { gone[del] = 1; anyGone = 1; }
                }
            }
            s1 = anchor; s2 = anchor + 1;
        } else { s1 = s2; s2 = s3; }
        if (s2 < END) s3 = s2 + 1;
    }
    int n = END;
    if (anyGone) {
        n = 0;
        for (int k = 0; k < END; k++) if (!gone[k]) p[n++] = p[k];
    }
    /* first and last fragments: only dropped when they abut their neighbour (AlignHelpers.c:154-192) */
    if (n == 0) return 0;
    if (fc_qlen(&p[0]) < P->wordLen && n > 1) {
        const ya_frag *f1 = &p[0], *f2 = &p[1];
        const int qGap = (int)fc_gap(f1->endQueryOff, f2->startQueryOff), rGap = (int)fc_gap_u(fc_ero(f1), f2->startRefOff);
        if ((qGap == 0 && rGap <= 2 * P->bandWidth) || (rGap == 0 && qGap <= 2 * P->bandWidth)) {
            for (int k = 1; k < n; k++) p[k - 1] = p[k];
            n--;
        }
    }
    if (fc_qlen(&p[n - 1]) < P->wordLen) {
        if (n == 1) return n;
        const ya_frag *f1 = &p[n - 2], *f2 = &p[n - 1];
        const int qGap = (int)fc_gap(f1->endQueryOff, f2->startQueryOff), rGap = (int)fc_gap_u(fc_ero(f1), f2->startRefOff);
        if ((qGap == 0 && rGap <= 2 * P->bandWidth) || (rGap == 0 && qGap <= 2 * P->bandWidth)) n--;
    }
    return n;
}

/* buildBestClumpFromFragmentRange (GraphPath.cpp:161-270) over the unused fragments frags[lo..hi]: best path by
 * the reference's scores and tie rules, fragments inserted from the path's end to its start with insertFragment's
 * overlap chops (AlignHelpers.c:60-90; a chop of the fragment being inserted is written back to frags[], as the
 * reference edits the array element), then minMatch test and clean-up.  The clump's fragments go to dst[] in
 * query order; returns their number (0: no clump) and the matched bases through *matched. */
FC_HD int fc_build_best(const fc_params *P, ya_frag *frags, int lo, int hi, const uint8_t *used,
                        fc_node *nodes, ya_frag *tmp, uint8_t *gone, ya_frag *dst, uint16_t *matched)
{
    int nc = 0;
    if (FC_LANE == 0) {
        for (int i = lo; i <= hi; i++) {
            if (used[i - lo]) continue;
            const ya_frag *f = &frags[i];
            fc_node n;
            n.prev = -1; n.pathLength = 1; n.frag = i; n.diag = fc_diag(f);
            n.nodeLength = (int16_t)f->refLen; n.bestScore = (int16_t)(n.nodeLength * P->MScore);
            n.SQO = f->startQueryOff; n.EQO = f->endQueryOff; n.pathSQO = n.SQO;
            nodes[nc++] = n;
        }
        if (nc) fc_sort_nodes(nodes, nc);
    }
    FC_BCAST(nc);
    FC_SYNC();
    *matched = 0;
    if (nc == 0) return 0;
    int bestScore = -(0x7fffff00), best = -1;
    const uint32_t maxGap = (uint32_t)P->maxGap;
    for (int i = 0; i < nc; i++) {
        const fc_node *L = &nodes[i];
        const int lSQO = L->SQO, lEQO = L->EQO;
        const uint32_t lSRO = L->diag + (uint32_t)lSQO, lERO = L->diag + (uint32_t)L->EQO;
        /* every j updates its own node only: the iterations are strided over the lanes (nodes behind i with the same start as
         * i sit right behind it in the sorted order, so "stop at the first such node coming from the end" is per lane) */
        for (int j = nc - 1 - FC_LANE; j > i; j -= FC_NLANES) {
            fc_node *R = &nodes[j];
            const int rSQO = R->SQO;
            if (rSQO == lSQO) break;
            const uint32_t diagGap = fc_absdiff(L->diag, R->diag);
            if (diagGap > maxGap) continue;
            const uint32_t rSRO = R->diag + (uint32_t)rSQO;
            if (lSRO >= rSRO) continue;
            const int desert = (int)fc_min_u(fc_gap(lEQO, rSQO), fc_gap_u(lERO, rSRO));
            if (desert > P->maxDesert) continue;
            const int maxOverlap = (int)fc_max_u(fc_overlap(lEQO, rSQO), fc_overlap_u(lERO, rSRO));
            const int newbases = R->nodeLength - maxOverlap;
            if (newbases < 1) continue;
            const int gapCost = diagGap > 0 ? -(P->GOCost + (int)diagGap * P->GECost) : 0;           /* calcGapCost */
            const int newScore = L->bestScore + newbases * P->MScore + gapCost;
            if (R->bestScore > newScore) continue;
            if (R->bestScore == newScore) {
                if (R->prev < 0) continue;
                const fc_node *Q = &nodes[R->prev];
                const int diagCompare = (int)(fc_absdiff(L->diag, R->diag) - fc_absdiff(Q->diag, R->diag));
                if (diagCompare > 0) continue;
                if (diagCompare == 0) {
                    const int gapCompare = (int)(fc_gap(L->EQO, R->SQO) - fc_gap(Q->EQO, R->SQO));
                    if (gapCompare > 0) continue;
                    if (gapCompare == 0 && L->pathSQO <= Q->pathSQO) continue;
                }
            }
            R->bestScore = (int16_t)newScore; R->prev = i; R->pathLength = (int16_t)(L->pathLength + 1); R->pathSQO = L->pathSQO;
        }
        FC_SYNC();
        if (L->bestScore < bestScore) continue;
        int take = L->bestScore > bestScore;
        if (!take) {                                                  /* GraphPath.cpp:88-94 */
            const fc_node *B = &nodes[best];
            take = (L->EQO != B->EQO) ? (L->EQO < B->EQO) : (L->pathSQO > B->pathSQO);
        }
        if (take) { best = i; bestScore = L->bestScore; }
    }
    /* the path from its end to its start (GraphPath.cpp:134-146); tmp[] holds it in insertion order, i.e. reversed */
    int result = 0, mbOut = 0;
    if (FC_LANE == 0) {
        int cnt = 0;
        uint16_t mb = 0;
        for (int k = best; k >= 0; k = nodes[k].prev) {
            ya_frag *f1 = &frags[nodes[k].frag];
            if (cnt > 0) {                                            /* insertFragment, AlignHelpers.c:60-90 */
                ya_frag *f2 = &tmp[cnt - 1];                          /* the clump's current first fragment */
                const int maxOverlap = (int)fc_max_u(fc_overlap(f1->endQueryOff, f2->startQueryOff), fc_overlap_u(fc_ero(f1), f2->startRefOff));
                if (maxOverlap > 0) {
                    const int l1 = fc_qlen(f1), l2 = fc_qlen(f2);
                    const int chop1 = (l1 != l2) ? (l1 < l2) : (cnt == 1);
                    if (chop1) { f1->endQueryOff = (uint16_t)(f1->endQueryOff - maxOverlap); f1->refLen = (uint16_t)(f1->refLen - maxOverlap); }
                    else { f2->startQueryOff = (uint16_t)(f2->startQueryOff + maxOverlap); f2->startRefOff += (uint32_t)maxOverlap;
                           f2->refLen = (uint16_t)(f2->refLen - maxOverlap); }
                }
            }
            mb = (uint16_t)(mb + f1->refLen);                         /* addFragment, AlignHelpers.c:48-56 */
            tmp[cnt] = *f1;
            tmp[cnt].hitCount = 0;
            cnt++;
        }
        if ((int)mb >= P->minMatch) {
            for (int k = 0; k < cnt; k++) dst[k] = tmp[cnt - 1 - k];
            mbOut = mb;
            result = fc_clean_up(P, dst, cnt, gone);
        }
    }
    FC_BCAST(result); FC_BCAST(mbOut);
    FC_SYNC();
    *matched = (uint16_t)mbOut;
    return result;
}

/* One strand.  frags[] is edited in place (overlap chops).  Returns the number of clumps written to out_clumps
 * (at most n); their fragments are appended to out_path (at most n in total; `first` is relative to out_path). */
FC_HD int fc_form_clumps(const fc_params *P, ya_frag *frags, const uint32_t *region, int n, int readLen,
                         fc_node *nodes, uint8_t *used, ya_frag *tmp, ya_frag *out_path, ya_clump_rec *out_clumps)
{
    int nClumps = 0;
    uint32_t nPath = 0;
    const int qSlots = readLen + 1;
    uint8_t *gone = used + n;                                         /* used[] is 2n bytes: flags + clean-up marks */
    int i = 0;
    while (i < n) {
        int j = i;
        while (j + 1 < n && region[j + 1] == region[i]) j++;
        if (j == i) {                                                 /* QueryMatch.c:281-290 */
            if ((int)frags[i].refLen >= P->minMatch) {
                if (FC_LANE == 0) {
                    out_path[nPath] = frags[i];
                    out_path[nPath].hitCount = 0;
                    out_clumps[nClumps].first = nPath; out_clumps[nClumps].n = 1; out_clumps[nClumps].matchedBases = frags[i].refLen;
                }
                nClumps++; nPath++;
            }
        } else {                                                      /* GraphPath.cpp:272-292 */
            const int m = j - i + 1;
            for (int k = FC_LANE; k < m; k += FC_NLANES) used[k] = 0;
            FC_SYNC();
            int unused = m;
            const int firstClumpOfRegion = nClumps;
            while (unused > 0) {
                uint16_t matched = 0;
                const int len = fc_build_best(P, frags, i, j, used, nodes, tmp, gone, out_path + nPath, &matched);
                if (len == 0) break;
                if (FC_LANE == 0) { out_clumps[nClumps].first = nPath; out_clumps[nClumps].n = (uint16_t)len; out_clumps[nClumps].matchedBases = matched; }
                nClumps++; nPath += (uint32_t)len;
                FC_SYNC();
                /* eliminateFragments, QueryMatch.c:201-215 (+ :177-197): a fragment stays only if one of its ends
                 * (minNonOverlap bases) is untouched by every clump cut from this region so far */
                const int minLeft = P->minNonOverlap - 1;
                int removed = 0;
                if (FC_LANE == 0) {
                    for (int k = i; k <= j; k++) {
                        if (used[k - i]) continue;
                        const int SQO = frags[k].startQueryOff, EQO = frags[k].endQueryOff;
                        int keep = 0;
                        if (EQO - SQO >= minLeft) {
                            int freeLo = 1, freeHi = 1;
                            for (int c = firstClumpOfRegion; c < nClumps; c++) {
                                const ya_frag *cf = &out_path[out_clumps[c].first], *cl = &out_path[out_clumps[c].first + out_clumps[c].n - 1];
                                const int sqo = cf->startQueryOff;
                                int hi = sqo + (int)(uint16_t)(1 + cl->endQueryOff - cf->startQueryOff) - 1;
                                if (hi > qSlots - 1) hi = qSlots - 1;
                                if (hi < sqo) continue;
                                if (sqo <= SQO + minLeft && SQO <= hi) freeLo = 0;
                                if (sqo <= EQO && EQO - minLeft <= hi) freeHi = 0;
                            }
                            keep = freeLo || freeHi;
                        }
                        if (!keep) { used[k - i] = 1; removed++; }
                    }
                }
                FC_BCAST(removed);
                FC_SYNC();
                unused -= removed;
            }
        }
        i = j + 1;
    }
    return nClumps;
}

#endif
