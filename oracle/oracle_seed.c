/* oracle/oracle_seed.c -- CPU restatement of yaha's seed lookup and fragment formation
 * (stages 1, 2a, 2b).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows: seed-lookup loop                     Query.c:296-297,341,365-412
 *          generateMatches4UNPto2Fast           Query.c:233-244
 *          findFragmentsSort                    QueryMatch.c:52-121 (+ QueryHeap.inl:70-73 key)
 *          findAlignableFragsForw, region loop  QueryMatch.c:146-158,224-303
 *
 * The reference merges per-k-mer hit lists with a binary heap; because all keys are
 * distinct, the emitted order is exactly ascending key order, so this restatement sorts.
 */
#include <stdlib.h>
#include "oracle.h"

uint32_t orc_seed_lookup(const ya_params *p, const uint32_t *so, const uint8_t *codes, int L,
                         uint32_t *sOffset, uint32_t *count)
{
    const int K = p->wordLen;
    const uint32_t mask = 0xFFFFFFFFu >> (32 - 2 * K);
    uint32_t total = 0;
    for (int qo = 0; qo + K <= L; qo++) {
        /* a k-mer is usable iff all K codes are plain bases (Query.c:236-240,373-388) */
        uint32_t h = 0; int ok = 1;
        for (int k = 0; k < K; k++) {
            uint8_t c = codes[qo + k];
            if (c > 3) { ok = 0; break; }
            h = (h << 2) + c;
        }
        h &= mask;
        sOffset[qo] = 0; count[qo] = 0;
        if (!ok) continue;
        uint32_t cnt = so[h + 1] - so[h];                       /* Query.c:391 */
        if (cnt <= (uint32_t)p->maxHits) {                      /* Query.c:392 */
            total += cnt; count[qo] = cnt; sOffset[qo] = so[h];
        }
    }
    return total;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

int orc_find_frags(const ya_params *p, const uint32_t *roa, size_t n_roa,
                   const uint32_t *sOffset, const uint32_t *count, int matchCount,
                   ya_frag *frags, int frag_cap)
{
    const int K = p->wordLen;
    size_t cap = 1024, n = 0;
    uint64_t *keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
#define PUSH(roff, qo) do { if (n == cap) { cap *= 2; keys = (uint64_t *)realloc(keys, cap * sizeof(uint64_t)); } \
        keys[n++] = ((uint64_t)(uint32_t)((roff) - (uint32_t)(qo)) << 32) + (uint32_t)(qo); } while (0)
    for (int qo = 0; qo < matchCount; qo++) {
        uint32_t c = count[qo];
        if (c == 0) continue;
        const uint32_t *list = roa + sOffset[qo];
        for (uint32_t t = 0; t < c; t++) PUSH(list[t], qo);
        /* QueryMatch.c:62-67: the pre-load loop for wrapped diagonals has no bound on
           newCount.  If every hit of this k-mer lies below qo, it keeps reading the ROA
           past the k-mer's own list, inserting each foreign offset as a hit at this qo,
           up to and including the first one that is >= qo. */
        if (list[c - 1] < (uint32_t)qo) {
            size_t at = (size_t)sOffset[qo] + c;
            while (at < n_roa) {
                uint32_t x = roa[at++];
                PUSH(x, qo);
                if (x >= (uint32_t)qo) break;
            }
        }
    }
#undef PUSH
    if (n == 0) { free(keys); return 0; }
    qsort(keys, n, sizeof(uint64_t), cmp_u64);

    /* coalesce runs on one diagonal with abutting/overlapping seeds (QueryMatch.c:99-115) */
    int nf = 0;
    uint32_t curDiag = (uint32_t)(keys[0] >> 32);
    uint32_t firstQO = (uint32_t)(keys[0] & 0xFFFF), endQO = firstQO + K;   /* one past last base */
    for (size_t i = 1; i <= n; i++) {
        uint32_t d = 0, qo = 0;
        int brk = 1;
        if (i < n) {
            d = (uint32_t)(keys[i] >> 32); qo = (uint32_t)(keys[i] & 0xFFFF);
            brk = (d != curDiag) || ((uint16_t)qo > (uint16_t)endQO);
        }
        if (brk) {
            if (nf >= frag_cap) { free(keys); return -1; }
            ya_frag *f = &frags[nf++];
            f->startQueryOff = (uint16_t)firstQO;
            f->startRefOff = curDiag + firstQO;
            f->endQueryOff = (uint16_t)(endQO - 1);
            f->refLen = (uint16_t)(1 + (int)f->endQueryOff - (int)f->startQueryOff);
            f->hitCount = 0;
            curDiag = d; firstQO = qo; endQO = qo + K;
        } else endQO = qo + K;
    }
    free(keys);
    return nf;
}

int orc_regions(const ya_params *p, const ya_frag *frags, int n, uint32_t *region, uint8_t *keep)
{
    int nreg = 0, i = 0;
    while (i < n) {
        /* grow the region while neighbouring diagonals differ by <= maxGap (QueryMatch.c:146-158) */
        int j = i;
        uint32_t prev = frags[i].startRefOff - frags[i].startQueryOff;
        while (j + 1 < n) {
            uint32_t d = frags[j + 1].startRefOff - frags[j + 1].startQueryOff;
            uint32_t diff = prev > d ? prev - d : d - prev;
            if (diff > (uint32_t)p->maxGap) break;
            prev = d; j++;
        }
        int members = j - i + 1;
        for (int k = i; k <= j; k++) {
            region[k] = (uint32_t)nreg;
            keep[k] = (members > 1) || (frags[k].refLen >= p->minMatch);   /* QueryMatch.c:281-290 */
        }
        nreg++;
        i = j + 1;
    }
    return nreg;
}
