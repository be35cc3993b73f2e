/* assemble_clumps.h -- second phase of alignClump plus the verdict of scoreClump for ONE clump whose first DP round
 * (gap fills and both end extensions, prepare_clumps.h) has been answered.
 *
 * Plain C99 over flat arrays, stated once like form_clumps.h / prepare_clumps.h.  Today it is compiled for the host
 * program (host/align.cpp: every golden SAM pins it); it reads only what ya_prepare_clumps and ya_sw_batch leave on
 * the device (path, gap records, plan, result records, run-length ops, genome, read codes) and writes one record
 * plus a run of ops, so the same source is the body of the device kernel of row N2 (one thread per clump).
 *
 * Follows: alignClump splice + collapseSFragments        AlignHelpers.c:251-300  (junction rule mergeEOLToBack, SW.cpp:207-261)
 *          extendClumpForwardReverseTemplated            AlignExtFrag.cpp:64-143 (first, not "careful", extension)
 *          mergeEOLToFront / mergeEOLToBack               SW.cpp:151-261
 *          scoreClump (walk, split decision, thresholds)  AlignHelpers.c:302-366
 *
 * The reference builds the op list in five steps: concatenate the pieces (seed fragment "nM", gap answer, seed
 * fragment ...), lengthen the first and last run by the perfect end extensions, put the backward extension in front,
 * the forward extension behind.  Equal codes coalesce at every junction between two lists, never inside a DP answer.
 * Here the runs are written front to back in one pass with the same junction rule.
 */
#ifndef YAHA_B200_ASSEMBLE_CLUMPS_H
#define YAHA_B200_ASSEMBLE_CLUMPS_H
#include "prepare_clumps.h"

typedef struct ac_params {
    int32_t  GOCost, GECost, RCost, MScore, minExtLength, minRawScore;
    uint32_t maxROff;
    double   minIdentity;             /* AlignmentArgs_t.minIdentity (a float) widened, as the comparison AlignHelpers.c:358 does */
} ac_params;

enum { YA_ASM_DROP = 0, YA_ASM_SCORED = 1, YA_ASM_SPLIT = 2 };

typedef struct ya_asm_rec {
    ya_frag  frag;                    /* the collapsed fragment after both extensions                                  */
    int32_t  score;                   /* its aligned score (SFragment_t.score after AlignHelpers.c:264)                */
    uint32_t n_ops;                   /* runs written                                                                  */
    uint16_t matchedBases, mismatchedBases, gapBases, totLength, totScore;   /* Clump_t fields, set when SCORED      */
    uint8_t  verdict;                 /* YA_ASM_SCORED: done; YA_ASM_DROP: below -M / -P; YA_ASM_SPLIT: splitClump has to look at it */
    uint8_t  pad;
    uint32_t ops_off;                 /* set by the caller: where the clump's runs were written (device: index into the batch's run array) */
} ya_asm_rec;

#if defined(__CUDA_ARCH__) || !defined(__GNUC__)
#define AC_TOUCH(P) ((void)0)
#else
#define AC_TOUCH(P) __builtin_prefetch(P)       /* host: the answers' runs are read right after their records */
#endif

/* Upper bound of the runs ac_assemble_clump writes for this clump. */
FC_HD uint32_t ac_ops_bound(int np, const ya_gap_rec *gaps, int ng, const ya_prep_rec *prep, const ya_dp_result *res, const ya_op *rops)
{
    uint32_t n = (uint32_t)np;
    for (int g = 0; g < ng; g++) {
        if (gaps[g].job == 0xFFFFFFFFu) { n += 1u; continue; }
        const ya_dp_result *r = &res[gaps[g].job];
        n += r->ops_n; AC_TOUCH(rops + r->ops_off);
    }
    if (prep->jobB != 0xFFFFFFFFu) { const ya_dp_result *r = &res[prep->jobB]; n += r->ops_n; AC_TOUCH(rops + r->ops_off); }
    if (prep->jobF != 0xFFFFFFFFu) { const ya_dp_result *r = &res[prep->jobF]; n += r->ops_n; AC_TOUCH(rops + r->ops_off); }
    return n;
}

#define AC_SLOT(CODE) (((CODE) >> 1) & 7)      /* 'M' 6, 'R' 1, 'I' 4, 'D' 2 */

#define AC_RUN(CODE, LEN, JUNCTION)                                                                     \
    do {                                                                                                \
        if ((JUNCTION) && n > 0 && out[n - 1].opcode == (uint8_t)(CODE))                                \
            out[n - 1].length = (uint16_t)(out[n - 1].length + (LEN));                                  \
        else { out[n].length = (uint16_t)(LEN); out[n].opcode = (uint8_t)(CODE); out[n].pad = 0; n++; } \
    } while (0)

/* p[0..np): the clump's seed fragments after phase 1; gaps[0..ng) and *prep: its records of phase 1; res / rops: the
 * answers of the DP round that ran phase 1's jobs (indexed by the records' job numbers).  Writes the runs to out
 * (room for ac_ops_bound) and *rec.  Returns 0, or -1 if the extension plan re-derived on the collapsed fragment
 * differs from *prep (cannot happen: the gap fills never move the outer ends; callers treat it as fatal). */
FC_HD int ac_assemble_clump(const ac_params *P, const uint8_t *bases, const uint8_t *q, int readLen,
                            const ya_frag *p, int np, const ya_gap_rec *gaps, int ng, const ya_prep_rec *prep,
                            const ya_dp_result *res, const ya_op *rops, ya_op *out, ya_asm_rec *rec)
{
    /* collapseSFragments: one fragment from the first start to the last end (AlignHelpers.c:283-300) */
    ya_frag f = p[0];
    f.endQueryOff = p[np - 1].endQueryOff;
    f.refLen = (uint16_t)(1 + fc_ero(&p[np - 1]) - f.startRefOff);
    /* perfect part of both end extensions (AlignExtFrag.cpp:76-107) */
    int mB = 0, mF = 0;
    int backLen = (int)(f.startQueryOff < f.startRefOff ? f.startQueryOff : f.startRefOff);
    if (backLen > 0) { mB = pc_perfect_backward(bases, q, &f, backLen); backLen -= mB; }
    const uint16_t qlen = (uint16_t)((readLen - 1) - f.endQueryOff);
    const uint32_t rlen = P->maxROff - fc_ero(&f);
    int forwLen = (int)(qlen < rlen ? qlen : rlen);
    if (forwLen > 0) { mF = pc_perfect_forward(bases, q, &f, forwLen); forwLen -= mF; }
    const int doB = backLen >= P->minExtLength, doF = forwLen >= P->minExtLength;
    if (doB != (prep->jobB != 0xFFFFFFFFu) || doF != (prep->jobF != 0xFFFFFFFFu) ||
        (doB && backLen != (int)prep->backLen) || (doF && forwLen != (int)prep->forwLen)) return -1;

    int score = (mB + mF) * P->MScore;
    uint32_t n = 0;
    if (doB) {                                                        /* AlignExtFrag.cpp:109-125; list goes in front */
        const ya_dp_result *r = &res[prep->jobB];
        if (r->score > 0) {
            const ya_op *o = rops + r->ops_off;
            for (uint32_t k = 0; k < r->ops_n; k++) AC_RUN(o[k].opcode, o[k].length, 0);
            score += r->score;
            f.startQueryOff = (uint16_t)(f.startQueryOff - r->addedQLen);
            f.startRefOff -= (uint32_t)r->addedRLen; f.refLen = (uint16_t)(f.refLen + r->addedRLen);
        }
    }
    int gi = 0;
    for (int it = 0; it < np; it++) {                                 /* AlignHelpers.c:241-261, 283-300 */
        const int ql = fc_qlen(&p[it]);
        AC_RUN('M', (uint16_t)ql + (it == 0 ? mB : 0), 1);
        score += P->MScore * ql;
        if (gi < ng && (int)gaps[gi].after == it) {
            const ya_gap_rec *g = &gaps[gi++];
            if (g->job != 0xFFFFFFFFu) {
                const ya_dp_result *r = &res[g->job];
                const ya_op *o = rops + r->ops_off;
                score += r->score;
                for (uint32_t k = 0; k < r->ops_n; k++) AC_RUN(o[k].opcode, o[k].length, k == 0);
            } else {
                score += g->score;
                AC_RUN(g->code, g->len, 1);
            }
        }
    }
    out[n - 1].length = (uint16_t)(out[n - 1].length + mF);
    if (doF) {                                                        /* AlignExtFrag.cpp:127-143; list goes behind */
        const ya_dp_result *r = &res[prep->jobF];
        if (r->score > 0) {
            const ya_op *o = rops + r->ops_off;
            for (uint32_t k = 0; k < r->ops_n; k++) AC_RUN(o[k].opcode, o[k].length, k == 0);
            score += r->score;
            f.endQueryOff = (uint16_t)(f.endQueryOff + r->addedQLen);
            f.refLen = (uint16_t)(f.refLen + r->addedRLen);
        }
    }
    rec->frag = f; rec->score = score; rec->n_ops = n; rec->pad = 0; rec->ops_off = 0;
    rec->matchedBases = rec->mismatchedBases = rec->gapBases = rec->totLength = rec->totScore = 0;

    /* scoreClump (AlignHelpers.c:302-366): running score over the runs; a clump whose score touches zero, reaches its
     * total before the end, or ends below its maximum has to be split */
    /* score of a run (AlignHelpers.c:312-328): M +MScore*len, R -RCost*len, I and D -(GOCost + GECost*len)
     * (M, R, I, D -- the only codes a run carries here -- fall into four different slots under (code >> 1) & 7, so the
     * per-code sums and the score a * len + b need no data-dependent branch) */
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mul[8] = {0, 0, 0, 0, 0, 0, 0, 0}, add[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    mul[AC_SLOT('M')] = P->MScore; mul[AC_SLOT('R')] = -P->RCost;
    mul[AC_SLOT('I')] = mul[AC_SLOT('D')] = -P->GECost; add[AC_SLOT('I')] = add[AC_SLOT('D')] = -P->GOCost;
    int AGS = 0, maxAGS = 0, split = 0;
    for (uint32_t k = 0; k < n; k++) {
        const int h = AC_SLOT(out[k].opcode), len = out[k].length;
        cnt[h] += len;
        AGS += mul[h] * len + add[h];
        if (AGS <= 0 || (AGS >= score && k != n - 1)) { split = 1; break; }
        if (AGS > maxAGS) maxAGS = AGS;
    }
    const int matches = cnt[AC_SLOT('M')], mism = cnt[AC_SLOT('R')], ins = cnt[AC_SLOT('I')], del = cnt[AC_SLOT('D')];
    if (!split && matches >= P->minRawScore && maxAGS > AGS) split = 1;
    if (split) { rec->verdict = YA_ASM_SPLIT; return 0; }
    rec->verdict = YA_ASM_DROP;
    if (matches < P->minRawScore) return 0;
    rec->matchedBases = (uint16_t)matches; rec->mismatchedBases = (uint16_t)mism; rec->gapBases = (uint16_t)(ins + del);
    rec->totLength = (uint16_t)(matches + mism + ins + del); rec->totScore = (uint16_t)AGS;
    const double percent = (double)rec->matchedBases / rec->totLength;
    if (percent < P->minIdentity) return 0;
    rec->verdict = YA_ASM_SCORED;
    return 0;
}

#endif
