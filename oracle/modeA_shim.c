/* oracle/modeA_shim.c -- "Mode A": the UNMODIFIED reference host calling libyaha_b200.so.
 *
 * TEST INFRASTRUCTURE (it links the reference's objects); it is also the worked example of
 * this is the binding a yaha maintainer would add.  Linked with
 *   -Wl,--wrap=findFragmentsSort,--wrap=findAGSAlignment,--wrap=findAGSAlignmentBanded,
 *       --wrap=findAGSForwardExtension,--wrap=findAGSBackwardExtension
 * so every cross-TU call through those seams (Math.h:401-408,554) lands here and is served by the
 * CUDA library through its C ABI, one job per call (slow by construction -- the batched driver in
 * yaha_b200/host is the throughput path).  The careful extension variants (SW.cpp:553-788) call the
 * DP template directly inside SW.cpp and therefore stay on the reference's CPU path in this mode.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Math.h"
#include "SW.inl"
#include "../include/yaha_b200.h"

static ya_ctx *ctx;
static char cur_id[MAX_QUERY_ID_LEN + 1];
static int cur_len = -1;
static ya_strand_frags strands[2];
static ya_frag *frags; static uint32_t *regions; static size_t frags_cap;
static int have_frags;

static void die(const char *what)
{
    fprintf(stderr, "modeA: %s: %s\n", what, ya_last_error(ctx));
    exit(1);
}

static void ensure_ctx(QueryState_t *QS)
{
    if (ctx) return;
    AlignmentArgs_t *A = QS->AAs;
    ya_params p;
    p.wordLen = A->wordLen; p.maxHits = A->maxHits; p.bandWidth = A->bandWidth; p.maxGap = A->maxGap; p.maxIntron = A->maxIntron;
    p.minMatch = A->minMatch; p.GOCost = A->GOCost; p.GECost = A->GECost; p.RCost = A->RCost; p.MScore = A->MScore;
    p.XCutoff = A->XCutoff; p.minExtLength = A->minExtLength;
    size_t n_so = ((size_t)1 << (2 * A->wordLen)) + 1;
    size_t n_roa = ((UINT *)A->startingOffs)[-1];            /* header word 3: totalMatches (Index.c:171-176) */
    size_t n_bytes = ((size_t)A->maxROff + 1) / 2 + 4;
    ctx = ya_open(0, &p, A->startingOffs, n_so, A->ROAPtr, n_roa, (const uint8_t *)A->basePtr, n_bytes, A->maxROff);
    if (!ctx) { fprintf(stderr, "modeA: ya_open failed: %s\n", ya_last_error(NULL)); exit(1); }
}

/* make the current read resident (one-read batch) */
static void ensure_read(QueryState_t *QS)
{
    ensure_ctx(QS);
    if (cur_len == QS->queryLen && memcmp(cur_id, QS->queryID, QS->queryIDLen) == 0 && cur_id[QS->queryIDLen] == 0) return;
    memcpy(cur_id, QS->queryID, QS->queryIDLen); cur_id[QS->queryIDLen] = 0; cur_len = QS->queryLen;
    uint64_t offs[2] = {0, QS->queryLen};
    ya_read_batch b = {1, (const uint8_t *)QS->forwardCodeBuf, offs};
    if (ya_reads_upload(ctx, &b) != YA_OK) die("ya_reads_upload");
    have_frags = 0;
}

int __wrap_findFragmentsSort(AlignmentArgs_t *AAs, QueryState_t *QS, int matchCount)
{
    (void)AAs; (void)matchCount;
    ensure_read(QS);
    if (!have_frags) {
        for (;;) {
            ya_frag_batch fb;
            memset(&fb, 0, sizeof fb);
            fb.frags_cap = frags_cap; fb.strands = strands; fb.frags = frags; fb.region = regions;
            int rc = ya_seed_frags(ctx, &fb);
            if (rc == YA_E_CAPACITY) {
                frags_cap = fb.frags_needed + 64;
                frags = realloc(frags, frags_cap * sizeof(ya_frag)); regions = realloc(regions, frags_cap * sizeof(uint32_t));
                continue;
            }
            if (rc != YA_OK) die("ya_seed_frags");
            break;
        }
        have_frags = 1;
    }
    /* hand the reference the surviving fragments of this strand; dropped ones are singleton regions
       below minMatch, which processFragmentsGapped would discard anyway (QueryMatch.c:281-290) */
    const ya_strand_frags *s = &strands[QS->reversed ? 1 : 0];
    for (uint32_t k = 0; k < s->n_frags; k++) {
        const ya_frag *f = &frags[s->first + k];
        Fragment_t *o = QS->fragArray + k;
        o->startRefOff = f->startRefOff; o->startQueryOff = f->startQueryOff; o->endQueryOff = f->endQueryOff; o->refLen = f->refLen;
    }
    return (int)s->n_frags;
}

static int run_job(QueryState_t *QS, int kind, ROFF rOff, int rLen, char *qStr, int qOff, int qLen,
                   ya_dp_result *res, ya_op *ops, size_t cap)
{
    ensure_read(QS);
    ya_dp_job j;
    j.rOff = rOff; j.read = 0; j.rLen = (uint16_t)rLen; j.qOff = (uint16_t)qOff; j.qLen = (uint16_t)qLen; j.kind = (uint8_t)kind;
    j.strand = (qStr == QS->reverseCodeBuf);
    size_t need = 0;
    if (ya_sw_batch(ctx, &j, 1, res, ops, cap, &need) != YA_OK) die("ya_sw_batch");
    return res->score;
}

static void append_ops(EditOpList_t *list, const ya_op *ops, uint32_t n)
{
    for (uint32_t k = 0; k < n; k++) addEditOpToBack(list, (EditOpCode)ops[k].opcode, ops[k].length);
}

static ya_op opbuf[70000];

int __wrap_findAGSAlignment(QueryState_t *QS, ROFF rOff, QOFF rLen, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list)
{
    ya_dp_result r;
    run_job(QS, YA_DP_FULL, rOff, rLen, qStr, qOff, qLen, &r, opbuf, 70000);
    append_ops(list, opbuf, r.ops_n);
    return r.score;
}

int __wrap_findAGSAlignmentBanded(QueryState_t *QS, ROFF rOff, QOFF rLen, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list)
{
    ya_dp_result r;
    run_job(QS, YA_DP_BANDED, rOff, rLen, qStr, qOff, qLen, &r, opbuf, 70000);
    append_ops(list, opbuf, r.ops_n);
    return r.score;
}

static int ext(QueryState_t *QS, int kind, ROFF rOff, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list, QOFF *aq, QOFF *ar)
{
    ya_dp_result r;
    run_job(QS, kind, rOff, 0, qStr, qOff, qLen, &r, opbuf, 70000);
    *aq = r.addedQLen; *ar = r.addedRLen;
    if (r.score <= 0) return 0;
    EditOpList_t tmp;
    initEditOpList(&tmp, QS);
    append_ops(&tmp, opbuf, r.ops_n);
    if (kind == YA_DP_EXT_BWD) mergeEOLToFront(list, &tmp); else mergeEOLToBack(list, &tmp);     /* SW.cpp:528-531 */
    return r.score;
}

int __wrap_findAGSForwardExtension(QueryState_t *QS, ROFF rOff, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list, QOFF *aq, QOFF *ar)
{
    return ext(QS, YA_DP_EXT_FWD, rOff, qStr, qOff, qLen, list, aq, ar);
}

int __wrap_findAGSBackwardExtension(QueryState_t *QS, ROFF rOff, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list, QOFF *aq, QOFF *ar)
{
    return ext(QS, YA_DP_EXT_BWD, rOff, qStr, qOff, qLen, list, aq, ar);
}
