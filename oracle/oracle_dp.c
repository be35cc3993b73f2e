/* oracle/oracle_dp.c -- CPU restatement of yaha's affine-gap DP family (stage 3).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Written from the algorithm's definition with
 * whole-matrix storage so that it is easy to audit; it is deliberately not fast.
 *
 * Follows: findAffineGapScore<banded,extension,reverse,XCutoff>   SW.cpp:798-1208
 *          findAGSAlignment / findAGSAlignmentBanded              SW.cpp:462-475
 *          findAGSExtension<reverse> (clamping, <=0 rule)         SW.cpp:479-533
 *          decompressRef<reverse>                                 SW.cpp:444-456
 *          getFrom4Code                                           Math.c:180-188
 *          extendFragment{Forward,Backward}ToStopPerfectly        AlignExtFrag.cpp:30-48
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define WORST (-(0x7fffff00))          /* SW.cpp:356 */

int orc_base(const uint8_t *bases, uint32_t off)
{
    uint8_t b = bases[off >> 1];
    return (off & 1) ? (b & 0xF) : (b >> 4);
}

/* Math.c:141-156 restated as rules instead of a 128-entry table. */
static int code_of_char(int c)
{
    static const char order[] = "TCAGNBDHKMRSVWXY";   /* index = code (Math.c:154) */
    if (c >= 'a' && c <= 'z') c -= 32;
    if (c == 'U') return 0;
    if (c < 'A' || c > 'Z') return 14;
    const char *p = strchr(order, c);
    return p ? (int)(p - order) : 14;
}

static int comp_code(int code)
{
    /* complement in the character domain, then re-encode (equals Math.c:155) */
    static const char comp[] = "AGTCNVHDMKYSBWXR";    /* complement char of code i */
    return code_of_char(comp[code]);
}

void orc_encode(const char *chars, int n, uint8_t *fwd, uint8_t *rev)
{
    for (int i = 0; i < n; i++) {
        int c = code_of_char((unsigned char)chars[i]);
        fwd[i] = (uint8_t)c;
        rev[n - 1 - i] = (uint8_t)comp_code(c);
    }
}

int orc_perfect(const uint8_t *bases, const uint8_t *codes, uint32_t rOff, int qOff, int len, int dir)
{
    int n = 0;
    while (n < len && codes[qOff + dir * n] == orc_base(bases, rOff + (uint32_t)(dir * n))) n++;
    return n;
}

typedef struct { uint8_t op; int len; } bp_t;          /* back pointer of one cell */

/* The DP proper.  q(i) for i in 1..qLen and r[0..rLen) are already oriented.
 * Returns score; for extensions also the argmax cell via *maxi,*maxj (band coordinates). */
static int dp_fill_and_trace(const ya_params *p, int banded, int extension, int reverse,
                             const uint8_t *qbase, int qLen, const uint8_t *r, int rLen,
                             int *addedQLen, int *addedRLen,
                             ya_op *ops, int ops_cap, int *n_ops, int64_t *cells)
{
    const int GOC = p->GOCost, GEC = p->GECost, RC = p->RCost, MS = p->MScore;
    int lb = 0, rb = 0, W;
    if (banded) {
        if (extension) lb = rb = 2 * p->bandWidth;
        else {
            lb = p->bandWidth + (qLen > rLen ? qLen - rLen : 0);
            rb = p->bandWidth + (rLen > qLen ? rLen - qLen : 0);
        }
        W = lb + rb + 1;
    } else W = rLen + 1;
    const int origin = banded ? lb : 0;
    const int S = W + 1;                                   /* row stride incl. sentinel col */

    bp_t *bp = (bp_t *)calloc((size_t)(qLen + 1) * S, sizeof(bp_t));
    int *V = (int *)malloc(sizeof(int) * 2 * S), *F = (int *)malloc(sizeof(int) * 2 * S),
        *I = (int *)malloc(sizeof(int) * 2 * S);
    int *Vp = V, *Vc = V + S, *Fp = F, *Fc = F + S, *Ip = I, *Ic = I + S;

    /* row 0: origin, then leading deletes to its right (SW.cpp:884-916) */
    for (int j = 0; j < S; j++) { Vp[j] = WORST; Fp[j] = WORST; Ip[j] = 0; }
    Vp[origin] = 0; Fp[origin] = 0; Ip[origin] = 0;
    bp[origin].op = 'U'; bp[origin].len = 0;
    for (int j = origin + 1, k = 1; j < W; j++, k++) {
        Vp[j] = -(GOC + k * GEC); Fp[j] = WORST; Ip[j] = 0;
        bp[j].op = 'D'; bp[j].len = k;
    }
    /* leading inserts: first column (full) or the anti-diagonal left of the origin (banded),
       SW.cpp:922-933 */
    {
        int last = banded ? lb : qLen;
        for (int i = 1; i <= last && i <= qLen; i++) {
            int c = banded ? lb - i : 0;
            bp[(size_t)i * S + c].op = 'I'; bp[(size_t)i * S + c].len = i;
        }
    }

    int maxScore = WORST, maxi = 0, maxj = 0, lastV = 0;
    int64_t ncell = 0;
    for (int i = 1; i <= qLen; i++) {
        int startCol, endCol, Vleft;
        for (int j = 0; j < S; j++) { Vc[j] = WORST; Fc[j] = WORST; Ic[j] = 0; }
        if (banded) {
            startCol = lb + 1 - i;
            if (startCol <= 0) { startCol = 0; Vleft = WORST; }
            else {
                Vleft = -(GOC + i * GEC);                  /* SW.cpp:981 */
                Vc[startCol - 1] = Vleft;                  /* boundary cell, next row's diag */
            }
            endCol = lb + rLen - i; if (endCol > W - 1) endCol = W - 1;
        } else {
            startCol = 1; endCol = W - 1;
            Vleft = -(GOC + i * GEC);                      /* SW.cpp:988 */
            Vc[0] = Vleft;
        }
        int Eleft = WORST, Dleft = 0;                      /* SW.cpp:965-966 */
        int rowMax = WORST;
        int qc = reverse ? qbase[1 - i] : qbase[i - 1];    /* SW.cpp:999 */
        for (int j = startCol; j <= endCol; j++) {
            ncell++;
            int dj = banded ? j : j - 1;                   /* diag predecessor column  */
            int uj = dj + 1;                               /* insert predecessor column */
            int rc = banded ? r[i - lb - 1 + j] : r[j - 1];
            int v; uint8_t op; int len = 1;
            if (qc == rc) { v = Vp[dj] + MS; op = 'M'; } else { v = Vp[dj] - RC; op = 'R'; }
            /* delete (gap in query, consumes reference): SW.cpp:1029-1041 */
            int CE = Eleft - GEC, NE = Vleft - (GOC + GEC);
            if (CE >= NE && Dleft + 1 <= p->maxIntron) { Eleft = CE; Dleft = Dleft + 1; }
            else                                       { Eleft = NE; Dleft = 1; }
            if (extension ? (Eleft >= v) : (Eleft > v)) { v = Eleft; op = 'D'; len = Dleft; }
            /* insert (gap in reference, consumes query): SW.cpp:1046-1063 */
            int CF = Fp[uj] - GEC, NF = Vp[uj] - (GOC + GEC), f, ins;
            if (CF >= NF && Ip[uj] + 1 <= p->maxGap) { f = CF; ins = Ip[uj] + 1; }
            else                                     { f = NF; ins = 1; }
            if (extension ? (f >= v) : (f > v)) { v = f; op = 'I'; len = ins; }
            Fc[j] = f; Ic[j] = ins; Vc[j] = v;
            bp[(size_t)i * S + j].op = op; bp[(size_t)i * S + j].len = len;
            if (v > rowMax) rowMax = v;
            if (extension && v > maxScore) { maxScore = v; maxi = i; maxj = j; }
            Vleft = v;
            lastV = v;
        }
        int *t;
        t = Vp; Vp = Vc; Vc = t; t = Fp; Fp = Fc; Fc = t; t = Ip; Ip = Ic; Ic = t;
        if (extension && rowMax < maxScore - p->XCutoff) break;   /* SW.cpp:1091 */
    }
    if (cells) *cells += ncell;

    int score = extension ? maxScore : lastV;
    int nout = 0;
    if (extension && score <= 0) { score = 0; goto done; }           /* SW.cpp:1102 */
    if (extension) {
        *addedQLen = maxi;
        *addedRLen = maxi + (maxj - 2 * p->bandWidth);                /* SW.cpp:1109-1110 */
    } else {
        maxi = qLen; maxj = banded ? rb : rLen;                       /* SW.cpp:867-878 */
    }
    /* traceback, SW.cpp:1138-1195: walk back-pointers, merging equal neighbours */
    {
        int y = maxi, x = maxj;
        uint8_t prev = bp[(size_t)y * S + x].op; int run = 0;
        for (;;) {
            bp_t c = bp[(size_t)y * S + x];
            if (c.op == 'U') break;
            int len = c.len;
            if (c.op == 'D')      { x -= len; }
            else if (c.op == 'I') { y -= len; if (banded) x += len; }
            else                  { y -= 1; if (!banded) x -= 1; len = 1; }
            if (c.op != prev) {
                if (nout < ops_cap) { ops[nout].opcode = prev; ops[nout].length = (uint16_t)run; ops[nout].pad = 0; }
                nout++; prev = c.op; run = len;
            } else run += len;
        }
        if (nout < ops_cap) { ops[nout].opcode = prev; ops[nout].length = (uint16_t)run; ops[nout].pad = 0; }
        nout++;
        /* walking order is end -> start.  Forward jobs are prepended (genome order = reversed
           walk); reverse jobs are appended, and because the reverse DP walks from the far
           end towards the anchor, appended walk order is already genome order. */
        if (!reverse && nout <= ops_cap)
            for (int a = 0, b = nout - 1; a < b; a++, b--) { ya_op t = ops[a]; ops[a] = ops[b]; ops[b] = t; }
    }
done:
    if (n_ops) *n_ops = nout;
    free(bp); free(V); free(F); free(I);
    return score;
}

int orc_dp(const ya_params *p, const uint8_t *bases, uint32_t maxROff, const uint8_t *codes,
           int kind, uint32_t rOff, int rLen, int qOff, int qLen,
           int *addedQLen, int *addedRLen, ya_op *ops, int ops_cap, int *n_ops, int64_t *cells)
{
    int aq = 0, ar = 0, score;
    if (n_ops) *n_ops = 0;
    if (addedQLen) *addedQLen = 0;
    if (addedRLen) *addedRLen = 0;
    if (kind == YA_DP_FULL || kind == YA_DP_BANDED) {
        uint8_t *r = (uint8_t *)malloc((size_t)rLen + 1);
        for (int i = 0; i < rLen; i++) r[i] = (uint8_t)orc_base(bases, rOff + (uint32_t)i);
        score = dp_fill_and_trace(p, kind == YA_DP_BANDED, 0, 0, codes + qOff, qLen, r, rLen,
                                  &aq, &ar, ops, ops_cap, n_ops, cells);
        free(r);
        return score;
    }
    /* extensions: SW.cpp:479-533 */
    int reverse = (kind == YA_DP_EXT_BWD);
    if (qLen <= 0) return 0;
    int bw2 = 2 * p->bandWidth;
    uint32_t rl = (uint32_t)(qLen + bw2);
    if (reverse && rl > rOff) {
        rl = rOff + 1; qLen = (int)(rl - (uint32_t)bw2);
        if (qLen <= 0) return 0;
    }
    if (!reverse && rOff + rl > maxROff) {
        rl = maxROff - rOff; qLen = (int)(rl - (uint32_t)bw2);
        if (qLen <= 0) return 0;
    }
    uint8_t *r = (uint8_t *)malloc((size_t)rl + 1);
    for (uint32_t i = 0; i < rl; i++)
        r[i] = (uint8_t)orc_base(bases, reverse ? rOff - i : rOff + i);
    score = dp_fill_and_trace(p, 1, 1, reverse, codes + qOff, qLen, r, (int)rl,
                              &aq, &ar, ops, ops_cap, n_ops, cells);
    free(r);
    if (score <= 0) { if (n_ops) *n_ops = 0; return 0; }
    if (addedQLen) *addedQLen = aq;
    if (addedRLen) *addedRLen = ar;
    return score;
}
