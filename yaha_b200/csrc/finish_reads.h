/* finish_reads.h -- what follows a read's first DP round when none of its clumps has to be split: Optimal Query
 * Coverage over its scored clumps, the filter by similarity with mapping qualities, and the SAM text of the records
 * that survive (score/split walk + SAM record formatting).
 *
 * Plain C99 over flat arrays, stated once like form_clumps.h / prepare_clumps.h / assemble_clumps.h and compiled for
 * the device (finish.cu: one thread per read, ya_align_batch) and for the oracle-backed mock of the ABI
 * (tests/mock/mock_abi.c), where every golden SAM of the unmodified reference pins it.  The host program keeps its own
 * statement of the same functions over Clump objects (host/oqc.cpp, host/sam.cpp) for the reads this path hands back:
 * reads with a clump that splitClump has to look at, strands the device did not form, more scored clumps than
 * FR_MAX_NODES, -OQC N with two or more clumps (duplicate removal sorts with the C library's qsort), Blast8 output.
 *
 * Follows: clump graph nodes, key, RNG-tie-broken quicksort   GraphPath.cpp:298-459
 *          deleteSubsumedDups                                  GraphPath.cpp:461-517
 *          accurate overlap scoring, cached query lengths      GraphPath.cpp:704-878
 *          postFilterBySimilarity (best path search)           GraphPath.cpp:897-1086
 *          filterBySimilarity + mapping quality                GraphPath.cpp:526-692
 *          generateRandomSeed / xorshift                       QueryState.c:172-187, Math.c:274-284
 *          printClump (SAM)                                    AlignOutput.c:115-289
 *
 * Floating point: the reference compares ratios of small integers in double precision and rounds one product
 * (250.0 * ratio + 0.5).  Divisions are IEEE on both sides; the product-sum is written with FR_MUL / FR_ADD so that
 * the device does not contract it into a fused multiply-add (the reference's x86-64 build has none).  log10 of the
 * break-point distance (GraphPath.cpp:1018-1020) is NOT evaluated here: the caller tabulates, with the C library the
 * reference uses, the distances at which (int)(log10(d) * BPCost + 0.5) steps up (fr_params.bpp_dist).
 */
#ifndef YAHA_B200_FINISH_READS_H
#define YAHA_B200_FINISH_READS_H
#include "assemble_clumps.h"

#ifdef __CUDA_ARCH__
#define FR_MUL(a, b) __dmul_rn((a), (b))
#define FR_ADD(a, b) __dadd_rn((a), (b))
#else
#define FR_MUL(a, b) ((a) * (b))
#define FR_ADD(a, b) ((a) + (b))
#endif

#define FR_MAX_NODES 32
#define FR_WORST (-(0x7fffff00))

enum { FR_REVERSED = 1, FR_FORMED = 2, FR_ALIGNED = 4, FR_SCORED = 8, FR_SPLIT = 16, FR_PRIMARY = 32 };   /* FragsClumps.inl:221-226 */

typedef struct fr_params {
    int32_t GOCost, GECost, RCost, MScore;
    int32_t OQC, FBS, OQCMinNonOverlap, BPCost, maxBPLog;
    double  FBS_PSLength, FBS_PSScore;          /* AlignmentArgs_t floats, widened as the reference's comparisons do */
    int32_t hardClip, fastq;
    int32_t n_seq;
    const uint32_t *seq_start, *seq_len;        /* BaseSequence_t.startingOffset / .length (BaseSeq.c:115-119) */
    const uint32_t *seq_name_off;               /* n_seq + 1 offsets into seq_names */
    const char     *seq_names;
    int32_t bpp_base, n_bpp;                    /* break-point penalty for a distance d > 10: bpp_base + #{k : bpp_dist[k] <= d} */
    const uint32_t *bpp_dist;
} fr_params;

/* one scored clump of the read, as assemble_clumps.h left it */
typedef struct fr_clump {
    const ya_asm_rec *rec;
    const ya_op      *ops;                      /* rec->n_ops runs in genome order */
    int32_t           reversed;
} fr_clump;

typedef struct fr_node {                        /* cGraphNode, GraphPath.cpp:299-324 */
    int32_t  prev;                              /* index of best predecessor in the node array, -1 none */
    int32_t  clump;                             /* index into the read's fr_clump array, -1 = dead */
    int16_t  bestScore, pathLength;
    uint32_t SRO, ERO;
    uint16_t SQO, EQO;                          /* plus-strand normalised */
    int16_t  nodeLength, nodeScore;
    uint16_t qLenInOQC;
    uint8_t  reversed, seqNum;
} fr_node;

/* one record to print */
typedef struct fr_out {
    int32_t  clump;                             /* index into the read's fr_clump array */
    uint16_t matchedPrimary, numSecondaries;
    uint8_t  status, mapQuality;
} fr_out;

typedef struct fr_rand { uint32_t s[5]; } fr_rand;

FC_HD uint32_t fr_rand_bits(fr_rand *r)                                  /* Math.c:274-284 */
{
    const uint32_t t = r->s[0] ^ (r->s[0] >> 7);
    r->s[0] = r->s[1]; r->s[1] = r->s[2]; r->s[2] = r->s[3]; r->s[3] = r->s[4];
    r->s[4] = (r->s[4] ^ (r->s[4] << 6)) ^ (t ^ (t << 13));
    return (r->s[1] + r->s[1] + 1) * r->s[4];
}

FC_HD void fr_seed_random(fr_rand *r, const uint8_t *fcode, int len)     /* generateRandomSeed, QueryState.c:172-187 */
{
    int q = 0;
    for (int i = 0; i < 5; i++) {
        uint32_t word = 0;
        for (int j = 0; j < 16; j++) { word = (word << 2) | (fcode[q] & 3u); if (++q >= len) q = 0; }
        r->s[i] = word;
    }
}

/* findBaseSequenceNum, BaseSeq.c:81-90: the reference walks the sequences in order and returns the first that holds the offset.
 * Sequences lie one behind the other in ascending order (Compress.c:199-218), so the last one starting at or below the offset is
 * the only candidate: same answer by bisection -- a reference of 10^5 contigs would otherwise cost 10^5 compares per record. */
FC_HD int fr_find_seq(const fr_params *P, uint32_t off)
{
    int lo = 0, hi = P->n_seq;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (P->seq_start[mid] <= off) lo = mid + 1; else hi = mid; }
    const int i = lo - 1;
    if (i >= 0 && off < P->seq_start[i] + P->seq_len[i]) return i;
    return -1;
}

FC_HD uint64_t fr_key(const fr_node *n)                                  /* GraphPath.cpp:377-380 */
{
    return (((((uint64_t)n->SQO) << 16) + ((uint16_t)-(int16_t)n->EQO)) << 16) + ((uint16_t)-n->nodeScore);
}

FC_HD int fr_less(const fr_node *a, const fr_node *b, fr_rand *rng)      /* :382-388 */
{
    const uint64_t k1 = fr_key(a), k2 = fr_key(b);
    if (k1 == k2) return (fr_rand_bits(rng) & 1u) != 0;
    return k1 < k2;
}

/* quickSort, GraphPath.cpp:427-453: same pivots, same swaps, same order of comparisons (the tie break draws from the
 * generator), the recursion (left part first) unrolled onto an explicit stack */
FC_HD void fr_quick_sort(fr_node *a, int n, fr_rand *rng)
{
    int stackL[2 * FR_MAX_NODES + 4], stackR[2 * FR_MAX_NODES + 4], sp = 0;
    stackL[sp] = 0; stackR[sp] = n - 1; sp++;
    while (sp > 0) {
        sp--;
        const int left = stackL[sp], right = stackR[sp];
        if (left >= right) continue;
        const int pivot = (left + right) / 2;
        fr_node t = a[pivot]; a[pivot] = a[right]; a[right] = t;
        int store = left;
        for (int i = left; i < right; i++)
            if (fr_less(&a[i], &a[right], rng)) { t = a[i]; a[i] = a[store]; a[store] = t; store++; }
        t = a[store]; a[store] = a[right]; a[right] = t;
        /* right part is pushed first so that the left part is sorted first, as the recursion does */
        stackL[sp] = store + 1; stackR[sp] = right; sp++;
        stackL[sp] = left; stackR[sp] = store - 1; sp++;
    }
}

FC_HD int fr_score_for_length(const fr_params *P, const fr_clump *c, int length, int forward)     /* GraphPath.cpp:705-732 */
{
    int QLen = 0, AGS = 0;
    const int n = (int)c->rec->n_ops;
    for (int k = forward ? 0 : n - 1; k >= 0 && k < n && QLen < length; k += forward ? 1 : -1) {
        const ya_op o = c->ops[k];
        int len = o.length;
        if (o.opcode == 'D') AGS -= (P->GOCost + P->GECost * len);
        else {
            if (QLen + len > length) len = length - QLen;
            QLen += len;
            if (o.opcode == 'M') AGS += P->MScore * len;
            else if (o.opcode == 'R') AGS -= P->RCost * len;
            else if (o.opcode == 'I') AGS -= (P->GOCost + P->GECost * len);
        }
    }
    return AGS;
}

FC_HD int fr_accurate_overlap(const fr_params *P, const fr_clump *cl, fr_node *g, int left, int right, int overlap, int *rightBest)
{                                                                        /* GraphPath.cpp:744-800 */
    const fr_node *R = &g[right];
    const int rightScore = fr_score_for_length(P, &cl[R->clump], overlap, R->reversed ? 0 : 1);
    int pathScore = 0, remaining = overlap, cur = left;
    for (;;) {
        const fr_node *C = &g[cur];
        const int take = remaining < (int)C->qLenInOQC ? remaining : (int)C->qLenInOQC;
        remaining -= take;
        pathScore += fr_score_for_length(P, &cl[C->clump], take, C->reversed ? 1 : 0);
        if (remaining <= 0) break;
        cur = C->prev;
    }
    if (pathScore > rightScore) { *rightBest = 0; return rightScore; }
    *rightBest = 1;
    return pathScore;
}

FC_HD void fr_cache_qlen_reverse(fr_node *g, int left, int right, int overlap, int rightBest)      /* :802-826 */
{
    fr_node *R = &g[right];
    if (rightBest) {
        R->qLenInOQC = (uint16_t)(1 + R->EQO - R->SQO);
        int remaining = overlap, cur = left;
        for (;;) {
            fr_node *C = &g[cur];
            const int take = remaining < (int)C->qLenInOQC ? remaining : (int)C->qLenInOQC;
            C->qLenInOQC = (uint16_t)(C->qLenInOQC - take);
            remaining -= take;
            if (remaining <= 0) break;
            cur = C->prev;
        }
    } else R->qLenInOQC = (uint16_t)((1 + R->EQO - R->SQO) - overlap);
}

/* cacheQlenPath, GraphPath.cpp:841-867: recursion towards the start of the path, then the cached lengths are set
 * from the start of the path to `right` -- here the chain is collected first and walked back */
FC_HD void fr_cache_qlen_path(const fr_params *P, const fr_clump *cl, fr_node *g, int right)
{
    int chain[FR_MAX_NODES], n = 0;
    for (int p = right; p >= 0; p = g[p].prev) chain[n++] = p;
    for (int k = n - 1; k >= 0; k--) {
        fr_node *R = &g[chain[k]];
        const int qLen = 1 + R->EQO - R->SQO;
        if (R->prev < 0) { R->qLenInOQC = (uint16_t)qLen; continue; }
        const int left = chain[k + 1];
        const int overlap = (int)fc_overlap(g[left].EQO, R->SQO);
        if (overlap > 0) {
            int rb;
            fr_accurate_overlap(P, cl, g, left, chain[k], overlap, &rb);
            fr_cache_qlen_reverse(g, left, chain[k], overlap, rb);
        } else R->qLenInOQC = (uint16_t)qLen;
    }
}

FC_HD int fr_break_point_penalty(const fr_params *P, const fr_node *L, const fr_node *R)          /* GraphPath.cpp:1003-1026 */
{
    if (L->seqNum != R->seqNum) return P->maxBPLog * P->BPCost;
    uint32_t distance;
    if (L->SRO > R->ERO) distance = L->SRO - R->ERO;
    else if (R->SRO > L->ERO) distance = R->SRO - L->ERO;
    else distance = 0;
    if (distance <= 10) return P->BPCost;
    int lo = 0, hi = P->n_bpp;                                           /* #{k : bpp_dist[k] <= distance} */
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (P->bpp_dist[mid] <= distance) lo = mid + 1; else hi = mid; }
    return P->bpp_base + lo;
}

/* The whole tail of a read: cl[0..n) are its scored clumps in the order of the reference's clump list walk
 * (GraphPath.cpp:914-936: creation order, forward strand first).  Writes the records to print, in print order, to
 * outs (room for n) and *primaryCount; returns their number, or -1 when the read is not handled here. */
FC_HD int fr_finish_read(const fr_params *P, const uint8_t *fcode, int readLen, const fr_clump *cl, int n, fr_node *g, fr_out *outs,
                         int *primaryCount)
{
    *primaryCount = 0;
    if (n < 1) return 0;
    if (!P->OQC) {                                                       /* postFilterRemoveDups, GraphPath.cpp:1127-1174 */
        if (n >= 2) return -1;                                           /* (sorts with qsort: left to the host) */
        outs[0].clump = 0; outs[0].matchedPrimary = 0; outs[0].numSecondaries = 0;
        outs[0].status = (uint8_t)((cl[0].reversed ? FR_REVERSED : 0) | FR_ALIGNED | FR_SCORED); outs[0].mapQuality = 255;
        return 1;
    }
    if (n == 1) {                                                        /* GraphPath.cpp:903-912 */
        outs[0].clump = 0; outs[0].matchedPrimary = 1; outs[0].numSecondaries = 0;
        outs[0].status = (uint8_t)((cl[0].reversed ? FR_REVERSED : 0) | FR_ALIGNED | FR_SCORED | FR_PRIMARY); outs[0].mapQuality = 250;
        *primaryCount = 1;
        return 1;
    }
    if (n > FR_MAX_NODES) return -1;
    int cnt = 0;
    for (int k = 0; k < n; k++) {
        const ya_asm_rec *r = cl[k].rec;
        fr_node *nd = &g[cnt++];
        nd->prev = -1; nd->pathLength = 1; nd->clump = k;
        nd->bestScore = nd->nodeScore = (int16_t)(int)r->totScore;
        nd->nodeLength = (int16_t)r->totLength;
        const uint16_t sqo = r->frag.startQueryOff, eqo = r->frag.endQueryOff;
        if (cl[k].reversed) { nd->SQO = (uint16_t)((readLen - 1) - eqo); nd->EQO = (uint16_t)((readLen - 1) - sqo); }
        else { nd->SQO = sqo; nd->EQO = eqo; }
        nd->SRO = r->frag.startRefOff; nd->ERO = fc_ero(&r->frag);
        nd->reversed = (uint8_t)(cl[k].reversed != 0);
        nd->qLenInOQC = (uint16_t)(1 + eqo - sqo);
        nd->seqNum = (uint8_t)fr_find_seq(P, nd->SRO);
    }
    fr_rand rng;
    fr_seed_random(&rng, fcode, readLen);
    fr_quick_sort(g, cnt, &rng);

    /* deleteSubsumedDups, GraphPath.cpp:488-517 */
    int live = 0;
    for (int i = 0; i < cnt; i++) {
        if (g[i].clump < 0) continue;
        const fr_node cur = g[i];
        if (live != i) g[live] = g[i];
        live++;
        const int thr = cur.nodeScore / 8;
        for (int j = i + 1; j < cnt; j++) {
            fr_node *nx = &g[j];
            if (nx->clump < 0) continue;
            if (nx->EQO > cur.EQO) break;
            const int subsumed = (cur.EQO > nx->EQO && nx->nodeScore < thr);
            const int dup = (cur.SRO == nx->SRO && cur.ERO == nx->ERO && cur.reversed == nx->reversed && cur.SQO == nx->SQO && cur.EQO == nx->EQO);
            if (subsumed || dup) nx->clump = -1;
        }
    }
    cnt = live;

    /* best path, GraphPath.cpp:938-1079 */
    int bestScore = FR_WORST, best = -1;
    const int minNonOverlap = P->OQCMinNonOverlap;
    int startj = 1;
    for (int i = 0; i < cnt; i++) {
        fr_cache_qlen_path(P, cl, g, i);
        fr_node *Lf = &g[i];
        const int leftSQO = Lf->SQO, leftEQO = Lf->EQO;
        int foundstartj = 0;
        for (int j = startj; j < cnt; j++) {
            fr_node *R = &g[j];
            const int rightSQO = R->SQO;
            if ((rightSQO - leftSQO) >= minNonOverlap) {
                if (!foundstartj) { startj = j; foundstartj = 1; }
                const int rightEQO = R->EQO;
                if ((rightEQO - leftEQO) >= minNonOverlap) {
                    int16_t newScore = (int16_t)(Lf->bestScore + R->nodeScore);
                    if (!(R->bestScore > newScore)) {
                        newScore = (int16_t)(newScore - fr_break_point_penalty(P, Lf, R));
                        if (!(R->bestScore > newScore)) {
                            const int overlap = (int)fc_overlap(leftEQO, rightSQO);
                            int rightBest = 0, skip = 0;
                            if (overlap > 0) {
                                newScore = (int16_t)(newScore - fr_accurate_overlap(P, cl, g, i, j, overlap, &rightBest));
                                if (R->bestScore > newScore) skip = 1;
                            }
                            if (!skip && (R->bestScore < newScore || (R->prev >= 0 && Lf->pathLength < g[R->prev].pathLength))) {
                                if (overlap > 0) {                       /* cacheQlenInRightNode, :873-878 */
                                    const int qLen = 1 + R->EQO - R->SQO;
                                    R->qLenInOQC = (uint16_t)(rightBest ? qLen : qLen - overlap);
                                }
                                R->bestScore = newScore; R->prev = i; R->pathLength = (int16_t)(Lf->pathLength + 1);
                            }
                        }
                    }
                }
            }
            if (!foundstartj) startj = cnt;
        }
        if (Lf->bestScore < bestScore) continue;
        if (Lf->bestScore > bestScore || (best >= 0 && Lf->pathLength < g[best].pathLength)) { best = i; bestScore = Lf->bestScore; }
    }

    /* filterBySimilarity, GraphPath.cpp:571-692.  The reference's list after it: primaries from the end of the path to
     * its start, then the secondaries it keeps in node order; printClumps walks that list from its head, i.e. backwards. */
    const int primeCount = g[best].pathLength;
    int primIdx[FR_MAX_NODES];                                           /* node index of primary k (path order) */
    int alignedQLen[FR_MAX_NODES], numOutSec[FR_MAX_NODES];
    int16_t second[FR_MAX_NODES], third[FR_MAX_NODES];
    fr_out list[FR_MAX_NODES];                                           /* the reference's list, tail first */
    int nList = 0;
    {
        int pi = primeCount - 1;
        for (int p = best; p >= 0; p = g[p].prev) {
            primIdx[pi] = p;
            alignedQLen[pi] = 1 + g[p].EQO - g[p].SQO;
            second[pi] = 0; third[pi] = 0; numOutSec[pi] = 0;
            list[nList].clump = g[p].clump; list[nList].matchedPrimary = (uint16_t)(pi + 1); list[nList].numSecondaries = 0;
            list[nList].status = (uint8_t)((g[p].reversed ? FR_REVERSED : 0) | FR_ALIGNED | FR_SCORED | FR_PRIMARY);
            list[nList].mapQuality = 255;
            nList++;
            pi--;
        }
    }
    uint8_t isPrimary[FR_MAX_NODES];
    for (int i = 0; i < cnt; i++) isPrimary[i] = 0;
    for (int k = 0; k < primeCount; k++) isPrimary[primIdx[k]] = 1;
    const double targetOverlap = P->FBS_PSLength;
    for (int i = 0; i < cnt; i++) {
        if (isPrimary[i]) continue;
        const fr_node *cur = &g[i];
        const int curSQO = cur->SQO, curEQO = cur->EQO, curQLen = 1 + curEQO - curSQO;
        int maxOverlap = 0, maxIndex = 0;
        for (int k = 0; k < primeCount; k++) {
            const fr_node *pr = &g[primIdx[k]];
            const int ov = 1 + (curEQO < (int)pr->EQO ? curEQO : (int)pr->EQO) - (curSQO > (int)pr->SQO ? curSQO : (int)pr->SQO);
            if (ov > maxOverlap) { maxOverlap = ov; maxIndex = k; }
        }
        if (maxOverlap > 0) {
            if (cur->nodeScore > second[maxIndex]) { third[maxIndex] = second[maxIndex]; second[maxIndex] = cur->nodeScore; }
            else if (cur->nodeScore > third[maxIndex]) third[maxIndex] = cur->nodeScore;
            const fr_node *path = &g[primIdx[maxIndex]];
            if (((double)cur->nodeScore) / path->nodeScore >= P->FBS_PSScore) {
                const int ov = 1 + (curEQO < (int)path->EQO ? curEQO : (int)path->EQO) - (curSQO > (int)path->SQO ? curSQO : (int)path->SQO);
                const double ovD = ov;
                if (ovD / curQLen >= targetOverlap && ovD / alignedQLen[maxIndex] >= targetOverlap) {
                    numOutSec[maxIndex] += 1;
                    if (P->FBS) {
                        list[nList].clump = cur->clump; list[nList].matchedPrimary = (uint16_t)(maxIndex + 1); list[nList].numSecondaries = 0;
                        list[nList].status = (uint8_t)((cur->reversed ? FR_REVERSED : 0) | FR_ALIGNED | FR_SCORED);
                        list[nList].mapQuality = 255;
                        nList++;
                    }
                }
            }
        }
    }
    /* calcMQfromPAs, GraphPath.cpp:559-569 */
    for (int k = 0; k < primeCount; k++) {
        fr_out *o = &list[primeCount - 1 - k];                           /* primary k sits at list position primeCount-1-k */
        const double tot = (double)cl[o->clump].rec->totScore;
        if (second[k] == 0) o->mapQuality = 250;
        else {
            double a = tot - (double)second[k]; if (a < 0.0) a = 0.0;
            double ratio = a / tot;
            double b = tot - (double)third[k]; if (b < 0.0) b = 0.0;
            ratio = FR_MUL(ratio, FR_ADD(1.0, b / tot)) / 2.0;
            o->mapQuality = (uint8_t)FR_ADD(FR_MUL(250.0, ratio), 0.5);
        }
        o->numSecondaries = (uint16_t)numOutSec[k];
    }
    for (int k = 0; k < nList; k++) outs[k] = list[nList - 1 - k];       /* printClumps: from the list head */
    *primaryCount = primeCount;
    return nList;
}

/* ---- SAM record (printClump, AlignOutput.c:115-289).  One routine counts and writes: with w == NULL only the length
 * is returned, so the space a batch needs is known before a byte is written. ---- */
/* On the device a record is WRITTEN by a whole warp: all 32 lanes make the same call with the same arguments (the
 * control flow depends on the arguments only), single characters are stored by lane 0 and the long runs -- read bases,
 * qualities, reference bases of the MD tag -- are strided over the lanes.  Counting calls (w == NULL) store nothing and
 * may come from single threads. */
#ifdef __CUDA_ARCH__
#define FR_LANE  ((int)(threadIdx.x & 31u))
#define FR_LANES 32
#else
#define FR_LANE  0
#define FR_LANES 1
#endif
#define FR_PUTC(ch)   do { const char fr_ch_ = (char)(ch); if (w && FR_LANE == 0) w[n] = fr_ch_; n++; } while (0)

FC_HD size_t fr_put_uint(char *w, size_t n, uint32_t v)
{
    char buf[10]; int k = 0;
    do { buf[k++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (k) FR_PUTC(buf[--k]);
    return n;
}
FC_HD size_t fr_put_int(char *w, size_t n, int v)
{
    if (v < 0) { FR_PUTC('-'); return fr_put_uint(w, n, (uint32_t)(-(int64_t)v)); }
    return fr_put_uint(w, n, (uint32_t)v);
}
FC_HD size_t fr_put_str(char *w, size_t n, const char *s, size_t len)
{
    if (w) for (size_t k = (size_t)FR_LANE; k < len; k += FR_LANES) w[n + k] = s[k];
    return n + len;
}

FC_HD char fr_char_of_code(int c) { return "TCAGNBDHKMRSVWXY"[c & 15]; }   /* Math.c:154 */

/* id: the read's id (already cut to 200 characters, blanks replaced: Query.c:111-135); chars / qual: the read as it stands
 * in the file; rcode: its reverse-complement codes (the reverse strand is printed from codes, Query.c:164-167). */
FC_HD size_t fr_format_record(const fr_params *P, const uint8_t *bases, const char *id, int idLen, const char *chars, const char *qual,
                              const uint8_t *rcode, int readLen, const fr_clump *c, const fr_out *o, int primaryCount, char *w)
{
    size_t n = 0;
    const ya_asm_rec *r = c->rec;
    const ya_frag *f = &r->frag;
    uint32_t sStart = f->startRefOff;
    const uint32_t sEndAbs = fc_ero(f);
    const int si = fr_find_seq(P, sStart);
    if (si < 0 || sEndAbs >= P->seq_start[si] + P->seq_len[si]) return 0;     /* AlignOutput.c:129-136 */
    sStart -= P->seq_start[si];
    const int rev = (o->status & FR_REVERSED) != 0;
    n = fr_put_str(w, n, id, (size_t)idLen);
    if (rev) n = fr_put_str(w, n, "\t16\t", 4); else n = fr_put_str(w, n, "\t0\t", 3);
    n = fr_put_str(w, n, P->seq_names + P->seq_name_off[si], P->seq_name_off[si + 1] - P->seq_name_off[si]);
    FR_PUTC('\t'); n = fr_put_uint(w, n, sStart + 1); FR_PUTC('\t'); n = fr_put_uint(w, n, (uint32_t)o->mapQuality); FR_PUTC('\t');
    /* CIGAR: clips (AlignOutput.c:154-166), M and R merged */
    const char clipCh = P->hardClip ? 'H' : 'S';
    const int clipFront = f->startQueryOff, clipBack = readLen - 1 - f->endQueryOff;
    if (clipFront > 0) { n = fr_put_int(w, n, clipFront); FR_PUTC(clipCh); }
    int matches = 0;
    for (uint32_t k = 0; k < r->n_ops; k++) {
        const ya_op op = c->ops[k];
        if (op.opcode == 'M' || op.opcode == 'R') { matches += op.length; continue; }
        if (matches > 0) { n = fr_put_int(w, n, matches); FR_PUTC('M'); matches = 0; }
        n = fr_put_int(w, n, (int)op.length); FR_PUTC(op.opcode);
    }
    if (matches > 0) { n = fr_put_int(w, n, matches); FR_PUTC('M'); }
    if (clipBack > 0) { n = fr_put_int(w, n, clipBack); FR_PUTC(clipCh); }
    n = fr_put_str(w, n, "\t*\t0\t0\t", 7);
    int qs = 0, qe = readLen - 1;
    if (P->hardClip) { qs = f->startQueryOff; qe = f->endQueryOff; }
    if (qe >= qs) {
        if (w) {
            if (rev) for (int i = qs + FR_LANE; i <= qe; i += FR_LANES) w[n + (size_t)(i - qs)] = fr_char_of_code(rcode[i]);
            else for (int i = qs + FR_LANE; i <= qe; i += FR_LANES) w[n + (size_t)(i - qs)] = chars[i];
        }
        n += (size_t)(qe - qs + 1);
    }
    FR_PUTC('\t');
    if (P->fastq) {
        if (qe >= qs) {
            if (w) {
                if (rev) for (int i = qe - FR_LANE; i >= qs; i -= FR_LANES) w[n + (size_t)(qe - i)] = qual[i];
                else for (int i = qs + FR_LANE; i <= qe; i += FR_LANES) w[n + (size_t)(i - qs)] = qual[i];
            }
            n += (size_t)(qe - qs + 1);
        }
    } else FR_PUTC('*');
    n = fr_put_str(w, n, "\tAS:i:", 6); n = fr_put_int(w, n, (int)r->totScore);
    n = fr_put_str(w, n, "\tNM:i:", 6); n = fr_put_int(w, n, (int)r->gapBases + (int)r->mismatchedBases);
    n = fr_put_str(w, n, "\tMD:Z:", 6);
    matches = 0;
    int prev = (clipFront > 0) ? clipCh : 'U';
    uint32_t ro = f->startRefOff;
    for (uint32_t k = 0; k < r->n_ops; k++) {
        const ya_op op = c->ops[k];
        if (op.opcode == 'M') { matches += op.length; ro += op.length; }
        else if (op.opcode == 'R') {
            if (matches > 0) { n = fr_put_int(w, n, matches); matches = 0; }
            if (prev == 'D') FR_PUTC('0');
            if (w) for (int i = FR_LANE; i < (int)op.length; i += FR_LANES) w[n + (size_t)i] = fr_char_of_code(pc_base(bases, ro + (uint32_t)i));
            n += op.length;
            ro += op.length;
        } else if (op.opcode == 'D') {
            if (matches > 0) { n = fr_put_int(w, n, matches); matches = 0; }
            FR_PUTC('^');
            if (w) for (int i = FR_LANE; i < (int)op.length; i += FR_LANES) w[n + (size_t)i] = fr_char_of_code(pc_base(bases, ro + (uint32_t)i));
            n += op.length;
            ro += op.length;
        }
        prev = op.opcode;
    }
    if (matches > 0) n = fr_put_int(w, n, matches);
    n = fr_put_str(w, n, "\tYF:H:", 6);
    FR_PUTC("0123456789ABCDEF"[(o->status >> 4) & 15]); FR_PUTC("0123456789ABCDEF"[o->status & 15]);
    if (P->OQC) {
        n = fr_put_str(w, n, "\tYI:i:", 6); n = fr_put_int(w, n, (int)o->matchedPrimary);
        n = fr_put_str(w, n, "\tYP:i:", 6); n = fr_put_int(w, n, primaryCount);
        if (o->status & FR_PRIMARY) { n = fr_put_str(w, n, "\tYS:i:", 6); n = fr_put_int(w, n, (int)o->numSecondaries); }
    }
    FR_PUTC('\n');
    return n;
}

#endif
