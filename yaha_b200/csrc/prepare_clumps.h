/* prepare_clumps.h -- first phase of alignClump for the clumps of one strand:
 * perfect extensions between neighbouring seed fragments, classification of every gap (closed form or DP job),
 * and the plan of the two end extensions.  Plain C99, compiled for the device (clumps.cu: ya_prepare_clumps, one
 * thread per strand), for the host program (host/align.cpp) and for the mock of the ABI -- one statement, three users.
 *
 * Follows: alignClump phase 1                          AlignHelpers.c:205-261
 *          extendFragment{Forward,Backward}ToStopPerfectly   AlignExtFrag.cpp:30-48
 *          makeAndAlignSFragmentToFillGap (dispatch)    AlignExtFrag.cpp:164-234
 *          extendClumpForwardReverseTemplated (perfect part, lengths)   AlignExtFrag.cpp:64-107
 */
#ifndef YAHA_B200_PREPARE_CLUMPS_H
#define YAHA_B200_PREPARE_CLUMPS_H
#include "form_clumps.h"

typedef struct pc_params {
    int32_t bandWidth, GOCost, GECost, RCost, MScore, minExtLength;
    uint32_t maxROff;
} pc_params;

FC_HD int pc_base(const uint8_t *bases, uint32_t off) { const uint8_t b = bases[off >> 1]; return (off & 1) ? (b & 15) : (b >> 4); }   /* Math.c:180-188 */

FC_HD int pc_perfect_forward(const uint8_t *bases, const uint8_t *q, ya_frag *f, int len)      /* AlignExtFrag.cpp:30-38 */
{
    const uint16_t qOff = (uint16_t)(f->endQueryOff + 1);
    const uint32_t rOff = fc_ero(f) + 1;
    int n = 0;
    while (n < len && q[qOff + n] == pc_base(bases, rOff + (uint32_t)n)) n++;
    if (n > 0) { f->endQueryOff = (uint16_t)(f->endQueryOff + n); f->refLen = (uint16_t)(f->refLen + n); }
    return n;
}

FC_HD int pc_perfect_backward(const uint8_t *bases, const uint8_t *q, ya_frag *f, int len)     /* AlignExtFrag.cpp:40-48 */
{
    const uint16_t qOff = (uint16_t)(f->startQueryOff - 1);
    const uint32_t rOff = f->startRefOff - 1;
    int n = 0;
    while (n < len && q[qOff - n] == pc_base(bases, rOff - (uint32_t)n)) n++;
    if (n > 0) { f->startQueryOff = (uint16_t)(f->startQueryOff - n); f->startRefOff -= (uint32_t)n; f->refLen = (uint16_t)(f->refLen + n); }
    return n;
}

/* One clump: p[0..np) are its seed fragments in query order (edited: perfect extensions between neighbours).
 * Gaps go to gaps[] (at most np - 1), DP jobs to the job list through next_job (an index allocator shared by the
 * whole batch: an atomic counter on the device).  `read` / `strand` address the uploaded batch.  Returns the
 * number of gap records; *prep receives the extension plan. */
FC_HD int pc_prepare_clump(const pc_params *P, const uint8_t *bases, const uint8_t *q, int readLen, uint32_t read, int strand,
                           ya_frag *p, int np, ya_gap_rec *gaps, ya_dp_job *jobs, uint32_t *next_job, uint32_t jobs_cap,
                           ya_prep_rec *prep)
{
    for (int k = 1; k < np; k++) {                                    /* AlignHelpers.c:226-237 */
        ya_frag *l = &p[k - 1], *r = &p[k];
        int gap = (int)fc_min_u(fc_gap(l->endQueryOff, r->startQueryOff), fc_gap_u(fc_ero(l), r->startRefOff));
        gap -= pc_perfect_backward(bases, q, r, gap);
        gap -= pc_perfect_forward(bases, q, l, gap);
    }
    int ng = 0;
    for (int a = 0; a + 1 < np; a++) {                                /* AlignHelpers.c:251-261 + AlignExtFrag.cpp:164-234 */
        const ya_frag *f1 = &p[a], *f2 = &p[a + 1];
        const uint16_t qGap = (uint16_t)fc_gap(f1->endQueryOff, f2->startQueryOff);
        const uint16_t rGap = (uint16_t)fc_gap_u(fc_ero(f1), f2->startRefOff);
        if (qGap == 0 && rGap == 0) continue;
        ya_gap_rec g;
        g.after = (uint16_t)a; g.job = 0xFFFFFFFFu; g.score = 0; g.len = 0; g.code = 0; g.pad = 0; g.pad2 = 0;
        if (qGap == 0) { g.code = 'D'; g.len = rGap; g.score = -(P->GOCost + rGap * P->GECost); }
        else if (rGap == 0) { g.code = 'I'; g.len = qGap; g.score = -(P->GOCost + qGap * P->GECost); }
        else if (rGap == 1 && qGap == 1) { g.code = 'R'; g.len = 1; g.score = -P->RCost; }
        else {
            const int lenDiff = (int)qGap > (int)rGap ? (int)qGap - (int)rGap : (int)rGap - (int)qGap;
            const int banded = lenDiff + P->bandWidth * 2 + 1 < (int)rGap;
#ifdef __CUDA_ARCH__
            const uint32_t j = atomicAdd(next_job, 1u);
#else
            const uint32_t j = (*next_job)++;
#endif
            if (j < jobs_cap) {
                ya_dp_job jb;
                jb.rOff = fc_ero(f1) + 1; jb.read = read; jb.rLen = rGap; jb.qOff = (uint16_t)(f1->endQueryOff + 1); jb.qLen = qGap;
                jb.kind = (uint8_t)(banded ? YA_DP_BANDED : YA_DP_FULL); jb.strand = (uint8_t)strand;
                jobs[j] = jb;
            }
            g.job = j;
        }
        gaps[ng++] = g;
    }
    /* the first extensions start from the outer ends of the first and last fragment; the gap fills never move those
     * ends, so the plan can be made now (perfect pre-extension evaluated on copies) -- AlignExtFrag.cpp:64-107 */
    ya_frag f0 = p[0], fn = p[np - 1];
    int backLen = (int)(f0.startQueryOff < f0.startRefOff ? f0.startQueryOff : f0.startRefOff);
    if (backLen > 0) backLen -= pc_perfect_backward(bases, q, &f0, backLen);
    const uint16_t qlen = (uint16_t)((readLen - 1) - fn.endQueryOff);
    const uint32_t rlen = P->maxROff - fc_ero(&fn);
    int forwLen = (int)(qlen < rlen ? qlen : rlen);
    if (forwLen > 0) forwLen -= pc_perfect_forward(bases, q, &fn, forwLen);
    prep->backLen = (uint16_t)backLen; prep->forwLen = (uint16_t)forwLen;
    prep->jobB = prep->jobF = 0xFFFFFFFFu;
    if (backLen >= P->minExtLength) {
#ifdef __CUDA_ARCH__
        const uint32_t j = atomicAdd(next_job, 1u);
#else
        const uint32_t j = (*next_job)++;
#endif
        if (j < jobs_cap) {
            ya_dp_job jb;
            jb.rOff = f0.startRefOff - 1; jb.read = read; jb.rLen = 0; jb.qOff = (uint16_t)(f0.startQueryOff - 1); jb.qLen = (uint16_t)backLen;
            jb.kind = YA_DP_EXT_BWD; jb.strand = (uint8_t)strand;
            jobs[j] = jb;
        }
        prep->jobB = j;
    }
    if (forwLen >= P->minExtLength) {
#ifdef __CUDA_ARCH__
        const uint32_t j = atomicAdd(next_job, 1u);
#else
        const uint32_t j = (*next_job)++;
#endif
        if (j < jobs_cap) {
            ya_dp_job jb;
            jb.rOff = fc_ero(&fn) + 1; jb.read = read; jb.rLen = 0; jb.qOff = (uint16_t)(fn.endQueryOff + 1); jb.qLen = (uint16_t)forwLen;
            jb.kind = YA_DP_EXT_FWD; jb.strand = (uint8_t)strand;
            jobs[j] = jb;
        }
        prep->jobF = j;
    }
    prep->n_gaps = (uint16_t)ng; prep->pad = 0;
    return ng;
}

#endif
