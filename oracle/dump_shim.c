/* oracle/dump_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Linked together with the UNMODIFIED reference objects using
 *   -Wl,--wrap=findAGSAlignment,--wrap=findAGSAlignmentBanded,
 *       --wrap=findAGSForwardExtension,--wrap=findAGSBackwardExtension,
 *       --wrap=findFragmentsSort,--wrap=processFragmentsGapped
 * so that every cross-TU call through one of the hot-path seams is
 * forwarded to the real reference function and its inputs/outputs are appended as one text
 * line to the file named by $YAHA_DUMP.  Compiled against the reference's own header
 * (-I/root/reference/src); no reference source is copied.
 *
 * Record formats (space separated):
 *  D <kind> <qid> <strand> <rOff> <rLen> <qOff> <qLen> <score> <addQ> <addR> <ops|->
 *      kind: F full-global (Math.h:401), B banded-global (:402), E fwd ext (:403), R bwd ext (:407)
 *      ops : run-length list produced by THIS call, genome order, e.g. 12M1R3M ("-" if none)
 *  S <qid> <strand> <matchCount> <n> qo:sOffset:count ...   (stage-1 output, count>0 only)
 *  G <qid> <strand> <fragCount> sro:sqo:eqo:rlen ...        (findFragmentsSort output)
 *  C <qid> <strand> <nclumps> { <rev> <nfrags> sro:sqo:eqo:rlen ... } ...  (clump list after
 *      processFragmentsGapped, head first)
 */
#include <stdio.h>
#include <stdlib.h>
#include "Math.h"
#include "SW.inl"

static FILE *dumpf;
static int dump_mask = 0xF;   /* bit0 D, bit1 S, bit2 G, bit3 C */

static FILE *df(void)
{
    if (!dumpf) {
        const char *p = getenv("YAHA_DUMP");
        const char *m = getenv("YAHA_DUMP_MASK");
        if (m) dump_mask = atoi(m);
        dumpf = fopen(p ? p : "/dev/null", "w");
        if (!dumpf) { perror("YAHA_DUMP"); exit(1); }
        setvbuf(dumpf, NULL, _IOFBF, 1 << 20);
    }
    return dumpf;
}

__attribute__((destructor)) static void closedump(void) { if (dumpf) fclose(dumpf); }

static void put_qid(FILE *f, QueryState_t *QS)
{
    fwrite(QS->queryID, 1, QS->queryIDLen, f);
}

static int strand_of(QueryState_t *QS, char *qStr)
{
    return qStr == QS->reverseCodeBuf ? 1 : 0;
}

static void put_ops(FILE *f, EditOpList_t *list)
{
    if (EOLisEmpty(list)) { fputs(" -", f); return; }
    fputc(' ', f);
    forAllEditOpsInList(item, list) fprintf(f, "%u%c", item->length, item->opcode);
}

int __real_findAGSAlignment(QueryState_t *, ROFF, QOFF, char *, QOFF, QOFF, EditOpList_t *);
int __real_findAGSAlignmentBanded(QueryState_t *, ROFF, QOFF, char *, QOFF, QOFF, EditOpList_t *);
int __real_findAGSForwardExtension(QueryState_t *, ROFF, char *, QOFF, QOFF, EditOpList_t *, QOFF *, QOFF *);
int __real_findAGSBackwardExtension(QueryState_t *, ROFF, char *, QOFF, QOFF, EditOpList_t *, QOFF *, QOFF *);
int __real_findFragmentsSort(AlignmentArgs_t *, QueryState_t *, int);
void __real_processFragmentsGapped(AlignmentArgs_t *, QueryState_t *, int);

static int global_call(int banded, QueryState_t *QS, ROFF rOff, QOFF rLen, char *qStr, QOFF qOff, QOFF qLen,
                       EditOpList_t *list)
{
    int score = banded ? __real_findAGSAlignmentBanded(QS, rOff, rLen, qStr, qOff, qLen, list)
                       : __real_findAGSAlignment(QS, rOff, rLen, qStr, qOff, qLen, list);
    FILE *f = df();
    if (dump_mask & 1) {
        fprintf(f, "D %c ", banded ? 'B' : 'F');
        put_qid(f, QS);
        fprintf(f, " %d %u %u %u %u %d 0 0", strand_of(QS, qStr), rOff, rLen, qOff, qLen, score);
        put_ops(f, list);
        fputc('\n', f);
    }
    return score;
}

int __wrap_findAGSAlignment(QueryState_t *QS, ROFF rOff, QOFF rLen, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list)
{
    return global_call(0, QS, rOff, rLen, qStr, qOff, qLen, list);
}

int __wrap_findAGSAlignmentBanded(QueryState_t *QS, ROFF rOff, QOFF rLen, char *qStr, QOFF qOff, QOFF qLen, EditOpList_t *list)
{
    return global_call(1, QS, rOff, rLen, qStr, qOff, qLen, list);
}

static int ext_call(int reverse, QueryState_t *QS, ROFF rOff, char *qStr, QOFF qOff, QOFF qLen,
                    EditOpList_t *list, QOFF *addedQLen, QOFF *addedRLen)
{
    /* Run the real extension into an empty list so the ops of this call can be seen in
       isolation, then splice exactly as the reference does (SW.cpp:528-531). */
    EditOpList_t tmp;
    initEditOpList(&tmp, QS);
    int score = reverse ? __real_findAGSBackwardExtension(QS, rOff, qStr, qOff, qLen, &tmp, addedQLen, addedRLen)
                        : __real_findAGSForwardExtension(QS, rOff, qStr, qOff, qLen, &tmp, addedQLen, addedRLen);
    FILE *f = df();
    if (dump_mask & 1) {
        fprintf(f, "D %c ", reverse ? 'R' : 'E');
        put_qid(f, QS);
        fprintf(f, " %d %u 0 %u %u %d %u %u", strand_of(QS, qStr), rOff, qOff, qLen, score, *addedQLen, *addedRLen);
        put_ops(f, &tmp);
        fputc('\n', f);
    }
    if (reverse) mergeEOLToFront(list, &tmp);
    else         mergeEOLToBack(list, &tmp);
    return score;
}

int __wrap_findAGSForwardExtension(QueryState_t *QS, ROFF rOff, char *qStr, QOFF qOff, QOFF qLen,
                                   EditOpList_t *list, QOFF *aq, QOFF *ar)
{
    return ext_call(0, QS, rOff, qStr, qOff, qLen, list, aq, ar);
}

int __wrap_findAGSBackwardExtension(QueryState_t *QS, ROFF rOff, char *qStr, QOFF qOff, QOFF qLen,
                                    EditOpList_t *list, QOFF *aq, QOFF *ar)
{
    return ext_call(1, QS, rOff, qStr, qOff, qLen, list, aq, ar);
}

int __wrap_findFragmentsSort(AlignmentArgs_t *AAs, QueryState_t *QS, int matchCount)
{
    FILE *f = df();
    if (dump_mask & 2) {
        int n = 0;
        for (int i = 0; i < matchCount; i++) if (QS->offsetCounts[i].count) n++;
        fputs("S ", f);
        put_qid(f, QS);
        fprintf(f, " %d %d %d", QS->reversed, matchCount, n);
        for (int i = 0; i < matchCount; i++)
            if (QS->offsetCounts[i].count)
                fprintf(f, " %d:%u:%u", i, QS->offsetCounts[i].sOffset, QS->offsetCounts[i].count);
        fputc('\n', f);
    }
    int fragCount = __real_findFragmentsSort(AAs, QS, matchCount);
    if (dump_mask & 4) {
        fputs("G ", f);
        put_qid(f, QS);
        fprintf(f, " %d %d", QS->reversed, fragCount);
        for (int i = 0; i < fragCount; i++) {
            Fragment_t *fr = QS->fragArray + i;
            fprintf(f, " %u:%u:%u:%u", fr->startRefOff, fr->startQueryOff, fr->endQueryOff, fr->refLen);
        }
        fputc('\n', f);
    }
    return fragCount;
}

void __wrap_processFragmentsGapped(AlignmentArgs_t *AAs, QueryState_t *QS, int fragCount)
{
    __real_processFragmentsGapped(AAs, QS, fragCount);
    FILE *f = df();
    if (!(dump_mask & 8)) return;
    int n = 0;
    for (Clump_t *c = QS->clumps; c; c = c->next) n++;
    fputs("C ", f);
    put_qid(f, QS);
    fprintf(f, " %d %d", QS->reversed, n);
    for (Clump_t *c = QS->clumps; c; c = c->next) {
        int nf = 0;
        for (SFragment_t *s = c->SFragList.head; s; s = s->next) nf++;
        fprintf(f, " %d %d", c->status & 1, nf);
        for (SFragment_t *s = c->SFragList.head; s; s = s->next)
            fprintf(f, " %u:%u:%u:%u", s->frag.startRefOff, s->frag.startQueryOff, s->frag.endQueryOff, s->frag.refLen);
    }
    fputc('\n', f);
}
