// sam.cpp -- SAM / Blast8 record formatting (host-kept).
// Follows: outputFileHeader AlignOutput.c:30-111, printClump AlignOutput.c:115-321,
//          printClumps QueryMatch.c:333-344.  Records of one read are formatted into the read's
// own buffer so that the writer can emit reads in input order regardless of completion order.
#include <string.h>
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include "host.hpp"

namespace yh {

static void appendf(OutText &s, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
static void appendf(OutText &s, const char *fmt, ...)
{
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n > 0) s.append(buf, (size_t)std::min<int>(n, (int)sizeof buf - 1));
}

struct TwoDigits { char d[2]; uint8_t n; };
static const struct TwoDigitTable {
    TwoDigits t[100];
    TwoDigitTable()
    {
        for (int v = 0; v < 100; v++) {
            if (v < 10) { t[v].d[0] = (char)('0' + v); t[v].d[1] = '0'; t[v].n = 1; }
            else { t[v].d[0] = (char)('0' + v / 10); t[v].d[1] = (char)('0' + v % 10); t[v].n = 2; }
        }
    }
    const TwoDigits &operator[](unsigned v) const { return t[v]; }
} kTwoDigits;

void writeHeader(const Env &E, FILE *out)
{
    const Args &A = *E.A;
    if (!A.outputSAM) return;
    fprintf(out, "@HD\tVN:1.0\n");
    for (const BaseSeq &b : E.G->seqs) fprintf(out, "@SQ\tSN:%s\tLN:%u\n", b.name.c_str(), b.length);
    fprintf(out, "@PG\tID:YAHA\tVN:0.1.83\tCL:yaha");
    fprintf(out, " -q %s -x %s -os%c %s -t %d", A.qfile.c_str(), A.xfile.c_str(), A.hardClip ? 'h' : 's', A.ofile.c_str(), A.numThreads);
    fprintf(out, " -BW %d -G %d -H %d -M %d -MD %d -P %4.2f -X %d", A.bandWidth, A.maxGap, A.maxHits, A.minMatch, A.maxDesert,
            A.minIdentity, A.XCutoff);
    if (A.affineGapScoring) fprintf(out, " -AGS Y -GEC %d -GOC %d -MS %d -RC %d", A.GECost, A.GOCost, A.MScore, A.RCost);
    else fprintf(out, " -AGS N");
    if (A.OQC) {
        fprintf(out, " -OQC Y -BP %d -MGDP %d -MNO %d", A.BPCost, A.maxBPLog, A.OQCMinNonOverlap);
        if (A.FBS) fprintf(out, " -FBS Y -PRL %4.2f -PSS %4.2f", A.FBS_PSLength, A.FBS_PSScore);
        else fprintf(out, " -FBS N");
    } else fprintf(out, " -OQC N");
    fputc('\n', out);
}

static void formatClump(const Env &E, ReadCtx &rc, Clump &c)
{
    const Args &A = *E.A;
    const Genome &G = *E.G;
    OutText &o = *rc.out;
    const Frag &f0 = c.sf.front().frag, &fn = c.sf.back().frag;
    uint32_t sStart = f0.startRefOff, sEnd = fragERO(fn);
    int si = G.findSeq(sStart);
    if (si < 0 || sEnd >= G.seqs[si].start + G.seqs[si].length) return;        // AlignOutput.c:129-136
    const BaseSeq &BS = G.seqs[si];
    sStart -= BS.start; sEnd -= BS.start;
    const std::string &q = rc.chars(c.reversed());
    const int L = rc.read->len();
    if (A.outputSAM) {
        OpList &list = c.ops;
        int clip = L - 1 - f0.endQueryOff;
        if (clip > 0) list.pushBack(A.hardClip ? 'H' : 'S', clip);
        clip = f0.startQueryOff;
        if (clip > 0) list.pushFront(A.hardClip ? 'H' : 'S', clip);
        // the record is written through a raw cursor into space reserved for its largest possible size
        // (CIGAR <= 12 characters per run, MD <= 2 per reference base + 11 per run)
        size_t refBases = 0;
        for (const Op &op : list.v) if (op.code != 'I') refBases += op.len;
        const size_t bound = 512 + rc.read->id.size() + BS.name.size() + 24 * list.v.size() + 2 * (size_t)L + 2 * refBases;
        char *const w0 = o.room(bound);
        char *w = w0;
        auto putS = [&](const char *z, size_t n) { memcpy(w, z, n); w += n; };
        auto putU = [&](unsigned v) {                                      // (CIGAR / MD numbers are mostly one or two digits:
            if (v < 100) { memcpy(w, kTwoDigits[v].d, 2); w += kTwoDigits[v].n; return; }   //  both bytes stored, cursor moved by 1 or 2 -- no branch on the value)
            char buf[12]; int n = 0;
            do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (n) *w++ = buf[--n];
        };
        auto putI = [&](int v) { if (v < 0) { *w++ = '-'; putU((unsigned)(-(long)v)); } else putU((unsigned)v); };
        putS(rc.read->id.data(), rc.read->id.size());
        if (c.reversed()) putS("\t16\t", 4); else putS("\t0\t", 3);
        putS(BS.name.data(), BS.name.size());
        *w++ = '\t'; putU(sStart + 1); *w++ = '\t'; putU((unsigned)c.mapQuality); *w++ = '\t';
        int matches = 0;
        for (const Op &op : list.v) {
            if (op.code == 'M' || op.code == 'R') { matches += op.len; continue; }
            if (matches > 0) { putI(matches); *w++ = 'M'; matches = 0; }
            putI((int)op.len); *w++ = op.code;
        }
        if (matches > 0) { putI(matches); *w++ = 'M'; }
        putS("\t*\t0\t0\t", 7);
        int qs = 0, qe = L - 1;
        if (A.hardClip) { qs = f0.startQueryOff; qe = fn.endQueryOff; }
        if (qe >= qs) putS(q.data() + qs, (size_t)(qe - qs + 1));
        *w++ = '\t';
        if (A.fastq) {
            const std::string &ql = rc.read->qual;
            if (c.reversed()) for (int i = qe; i >= qs; i--) *w++ = ql[(size_t)i];
            else if (qe >= qs) putS(ql.data() + qs, (size_t)(qe - qs + 1));
        } else *w++ = '*';
        putS("\tAS:i:", 6); putI((int)c.totScore);
        putS("\tNM:i:", 6); putI((int)c.gapBases + (int)c.mismatchedBases);
        putS("\tMD:Z:", 6);
        matches = 0;
        char prev = 'U';
        uint32_t ro = f0.startRefOff;
        for (const Op &op : list.v) {
            if (op.code == prev) {                                              // AlignOutput.c:232-240
                fprintf(stderr, "Two identical codes in a row in EditOpList\n%s\n", rc.read->id.c_str());
                abort();
            }
            if (op.code == 'M') { matches += op.len; ro += op.len; }
            else if (op.code == 'R') {
                if (matches > 0) { putI(matches); matches = 0; }
                if (prev == 'D') *w++ = '0';
                for (int i = 0; i < op.len; i++) *w++ = kCharOfCode[G.code(ro + (uint32_t)i)];
                ro += op.len;
            } else if (op.code == 'D') {
                if (matches > 0) { putI(matches); matches = 0; }
                *w++ = '^';
                for (int i = 0; i < op.len; i++) *w++ = kCharOfCode[G.code(ro + (uint32_t)i)];
                ro += op.len;
            }
            prev = op.code;
        }
        if (matches > 0) putI(matches);
        putS("\tYF:H:", 6);
        { static const char hex[] = "0123456789ABCDEF"; const unsigned st = (unsigned)c.status; *w++ = hex[(st >> 4) & 15]; *w++ = hex[st & 15]; }
        if (A.OQC) {
            putS("\tYI:i:", 6); putI((int)c.matchedPrimary);
            putS("\tYP:i:", 6); putI(rc.primaryCount);
            if (c.is(kPrimary)) { putS("\tYS:i:", 6); putI((int)c.numSecondaries); }
        }
        *w++ = '\n';
        if ((size_t)(w - w0) > bound) { fprintf(stderr, "yaha_b200: internal error: SAM record larger than its bound\n"); abort(); }
        o.commit((size_t)(w - w0));
    }
    if (A.outputBlast8) {                                                       // AlignOutput.c:307-318
        o.append(rc.read->id.data(), rc.read->id.size()); o.append("\t", 1); o.append(BS.name.data(), BS.name.size());
        appendf(o, "\t%4.2f\t%d\t%d\t%d", 0.8 * 100, (int)c.totLength, (int)c.mismatchedBases, (int)c.gapBases);
        if (c.reversed()) appendf(o, "\t%d\t%d\t%d\t%d\t%c", L - fn.endQueryOff, L - f0.startQueryOff, sEnd + 1, sStart + 1, '-');
        else appendf(o, "\t%d\t%d\t%d\t%d\t%c", f0.startQueryOff + 1, fn.endQueryOff + 1, sStart + 1, sEnd + 1, '+');
        appendf(o, "\t%d\t%d\t%4.2f\n", (int)c.totScore, L, ((double)c.matchedBases / L) * 100);
    }
}

void formatClumps(const Env &E, ReadCtx &rc)
{
    rc.outOff = rc.out->size();
    for (int k = (int)rc.clumps.size() - 1; k >= 0; k--) formatClump(E, rc, *rc.clumps[k]);   // from the list head
    rc.outLen = rc.out->size() - rc.outOff;
}

}  // namespace yh
