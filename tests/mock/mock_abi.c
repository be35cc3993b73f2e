/* tests/mock/mock_abi.c -- TEST INFRASTRUCTURE ONLY.
 * Implements the C ABI of include/yaha_b200.h on top of the CPU oracle (oracle/oracle_*.c) so
 * that the host program's logic (graph, split/score, OQC, SAM, fiber scheduler) can be tested
 * and run under sanitizers on machines without a GPU.  It is linked only into
 * tests/_build/yaha_host_mock; the product never sees it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../oracle/oracle.h"
#include "../../yaha_b200/csrc/form_clumps.h"
#include "../../yaha_b200/csrc/prepare_clumps.h"

struct ya_ctx {
    ya_params P;
    const uint32_t *so, *roa; size_t n_so, n_roa;
    const uint8_t *bases; size_t n_base_bytes; uint32_t maxROff;
    int n_reads; uint8_t *fwd, *rev; uint64_t *off;
    ya_counters ctr;
    ya_clump_batch last_clumps_v; int has_clumps;     /* descriptor of the last ya_form_clumps outputs (caller's buffers) */
    ya_frag_batch last_seed_v; int has_seed;          /* descriptor of the last ya_seed_frags outputs (caller's buffers) */
    ya_op *pending; size_t pendingN, pendingCap;     /* ops of the last ya_sw_batch (for ya_sw_fetch_ops) */
    char err[256];
};
static const uint8_t comp[16] = {2, 3, 0, 1, 4, 12, 7, 6, 9, 8, 15, 11, 5, 13, 14, 10};

/* MOCK_MEMO=1: answers are remembered by a hash of the request, so that repeated passes over the same
 * reads (-passes N -replay, one thread) cost no oracle time -- used to profile the host logic alone. */
typedef struct memo { uint64_t key; int kind; void *a; size_t na; void *b; size_t nb; void *c; size_t nc; size_t extra; struct memo *next; } memo;
static memo *memo_head;
static int memo_on(void) { static int v = -1; if (v < 0) v = getenv("MOCK_MEMO") != NULL; return v; }
static uint64_t fnv(const void *p, size_t n, uint64_t h)
{
    const uint8_t *b = p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static memo *memo_find(uint64_t key, int kind)
{
    for (memo *m = memo_head; m; m = m->next) if (m->key == key && m->kind == kind) return m;
    return NULL;
}
static void *dupmem(const void *p, size_t n) { void *q = malloc(n ? n : 1); memcpy(q, p, n); return q; }
static memo *memo_add(uint64_t key, int kind)
{
    memo *m = calloc(1, sizeof *m);
    m->key = key; m->kind = kind; m->next = memo_head; memo_head = m;
    return m;
}

ya_ctx *ya_open(int device, const ya_params *p, const uint32_t *so, size_t n_so, const uint32_t *roa, size_t n_roa,
                const uint8_t *bases, size_t n_base_bytes, uint32_t maxROff)
{
    ya_ctx *c = calloc(1, sizeof *c);
    c->P = *p; c->so = so; c->roa = roa; c->n_so = n_so; c->n_roa = n_roa; c->bases = bases; c->n_base_bytes = n_base_bytes;
    c->maxROff = maxROff;
    (void)device;
    return c;
}
ya_ctx *ya_open_peer(int device, const ya_ctx *src) { ya_ctx *c = malloc(sizeof *c); *c = *src; c->fwd = c->rev = NULL; c->off = NULL; c->n_reads = 0; (void)device; return c; }
ya_ctx *ya_open_shared(const ya_ctx *src) { return ya_open_peer(0, src); }
ya_ctx *ya_open_build(int d, const ya_params *p, const uint8_t *b, size_t n, const uint32_t *s, const uint32_t *l, int ns, uint32_t mh, uint32_t sk)
{ (void)d; (void)p; (void)b; (void)n; (void)s; (void)l; (void)ns; (void)mh; (void)sk; return NULL; }
int ya_index_sizes(const ya_ctx *c, size_t *a, size_t *b) { *a = c->n_so; *b = c->n_roa; return 0; }
int ya_index_download(ya_ctx *c, uint32_t *so, uint32_t *roa) { (void)c; (void)so; (void)roa; return YA_E_STATE; }
void ya_close(ya_ctx *c) { if (!c) return; free(c->fwd); free(c->rev); free(c->off); free(c->pending); free(c); }
const char *ya_last_error(const ya_ctx *c) { return c ? c->err : "mock"; }
int ya_set_params(ya_ctx *c, const ya_params *p) { c->P = *p; return 0; }
int ya_set_stream(ya_ctx *c, void *s) { (void)c; (void)s; return 0; }
int ya_get_counters(ya_ctx *c, ya_counters *o) { *o = c->ctr; memset(&c->ctr, 0, sizeof c->ctr); return 0; }
int ya_measure_int32_peak(ya_ctx *c, double *a, double *b) { (void)c; *a = *b = 0; return 0; }
int ya_measure_gather_peak(ya_ctx *c, double *a) { (void)c; *a = 0; return 0; }

void *ya_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void ya_host_free(void *p) { free(p); }

int ya_reads_upload(ya_ctx *c, const ya_read_batch *b)
{
    free(c->fwd); free(c->rev); free(c->off);
    c->n_reads = b->n_reads;
    uint64_t total = b->n_reads ? b->offsets[b->n_reads] : 0;
    c->fwd = malloc(total + 1); c->rev = malloc(total + 1); c->off = malloc((b->n_reads + 1) * sizeof(uint64_t));
    memcpy(c->off, b->offsets, (b->n_reads + 1) * sizeof(uint64_t));
    if (total) memcpy(c->fwd, b->codes, total);
    for (int r = 0; r < b->n_reads; r++) {
        uint64_t s = c->off[r], e = c->off[r + 1];
        for (uint64_t i = s; i < e; i++) c->rev[s + (e - 1 - i)] = comp[c->fwd[i] & 15];
    }
    return 0;
}

int ya_seed_frags(ya_ctx *c, ya_frag_batch *out)
{
    size_t n_out = 0;
    int overflow = 0;
    uint64_t key = 0;
    if (memo_on()) {
        key = fnv(c->off, (c->n_reads + 1) * sizeof(uint64_t), 14695981039346656037ull);
        key = fnv(c->fwd, c->n_reads ? c->off[c->n_reads] : 0, key);
        memo *m = memo_find(key, 1);
        if (m && m->extra <= out->frags_cap) {
            memcpy(out->strands, m->a, m->na); memcpy(out->frags, m->b, m->nb); memcpy(out->region, m->c, m->nc);
            out->n_frags = m->extra;
            c->last_seed_v = *out; c->has_seed = 1;
            return 0;
        }
    }
    c->has_seed = 0; c->has_clumps = 0;
    for (int seg = 0; seg < 2 * c->n_reads; seg++) {
        int r = seg >> 1;
        int L = (int)(c->off[r + 1] - c->off[r]);
        const uint8_t *codes = ((seg & 1) ? c->rev : c->fwd) + c->off[r];
        ya_strand_frags *s = &out->strands[seg];
        s->first = (uint32_t)n_out; s->n_frags = 0; s->n_frags_all = 0; s->total_hits = 0;
        int m = L - c->P.wordLen + 1;
        if (m <= 0) continue;
        uint32_t *soff = malloc(m * 4), *cnt = malloc(m * 4);
        uint32_t total = orc_seed_lookup(&c->P, c->so, codes, L, soff, cnt);
        s->total_hits = total;
        if (total) {
            int cap = (int)total + 4 * L + 64;
            ya_frag *fr = malloc((size_t)cap * sizeof(ya_frag));
            int nf;
            while ((nf = orc_find_frags(&c->P, c->roa, c->n_roa, soff, cnt, m, fr, cap)) < 0) { cap *= 2; fr = realloc(fr, (size_t)cap * sizeof(ya_frag)); }
            uint32_t *reg = malloc((nf + 1) * 4); uint8_t *keep = malloc(nf + 1);
            orc_regions(&c->P, fr, nf, reg, keep);
            s->n_frags_all = (uint32_t)nf;
            for (int k = 0; k < nf; k++) {
                if (!keep[k]) continue;
                if (n_out < out->frags_cap) { out->frags[n_out] = fr[k]; out->region[n_out] = reg[k]; } else overflow = 1;
                n_out++; s->n_frags++;
            }
            free(fr); free(reg); free(keep);
        }
        free(soff); free(cnt);
    }
    if (overflow) { out->frags_needed = n_out; return YA_E_CAPACITY; }
    out->n_frags = n_out;
    c->last_seed_v = *out; c->has_seed = 1;
    if (memo_on()) {
        memo *m = memo_add(key, 1);
        m->na = 2 * (size_t)c->n_reads * sizeof(ya_strand_frags); m->a = dupmem(out->strands, m->na);
        m->nb = n_out * sizeof(ya_frag); m->b = dupmem(out->frags, m->nb);
        m->nc = n_out * 4; m->c = dupmem(out->region, m->nc);
        m->extra = n_out;
    }
    return 0;
}

/* the same source the device kernel runs (yaha_b200/csrc/form_clumps.h), one strand after the other */
int ya_form_clumps(ya_ctx *c, ya_clump_batch *out)
{
    const ya_frag_batch *fb = &c->last_seed_v;
    out->n_clumps = out->n_path = 0;
    c->has_clumps = 0;
    if (!c->has_seed) return YA_E_STATE;
    if (fb->n_frags > out->cap) return YA_E_CAPACITY;
    fc_params P;
    P.wordLen = c->P.wordLen; P.maxGap = c->P.maxGap; P.maxDesert = out->maxDesert; P.minMatch = c->P.minMatch;
    P.minNonOverlap = out->minNonOverlap; P.bandWidth = c->P.bandWidth; P.GOCost = c->P.GOCost; P.GECost = c->P.GECost; P.MScore = c->P.MScore;
    for (int s = 0; s < 2 * c->n_reads; s++) {
        const ya_strand_frags *sf = &fb->strands[s];
        const uint32_t n = sf->n_frags, first = n ? sf->first : 0;
        out->clump_first[s] = first; out->clump_count[s] = 0;
        if (!n) continue;
        if (n > 192) { out->clump_count[s] = 0xFFFFFFFFu; continue; }   /* as the device kernel: left to the host (clumps.cu) */
        ya_frag *work = malloc(n * sizeof(ya_frag)), *tmp = malloc(n * sizeof(ya_frag));
        fc_node *nodes = malloc(n * sizeof(fc_node));
        uint8_t *used = malloc(2 * (size_t)n);
        memcpy(work, fb->frags + first, n * sizeof(ya_frag));
        const int L = (int)(c->off[(s >> 1) + 1] - c->off[s >> 1]);
        const int nc = fc_form_clumps(&P, work, fb->region + first, (int)n, L, nodes, used, tmp, out->path + first, out->clumps + first);
        for (int k = 0; k < nc; k++) { out->clumps[first + k].first += first; out->n_path += out->clumps[first + k].n; }
        out->clump_count[s] = (uint32_t)nc; out->n_clumps += (size_t)nc;
        free(work); free(tmp); free(nodes); free(used);
    }
    c->last_clumps_v = *out; c->has_clumps = 1;
    return 0;
}

int ya_prepare_clumps(ya_ctx *c, ya_prep_batch *out)
{
    const ya_clump_batch *cb = &c->last_clumps_v;
    out->n_jobs = 0;
    if (!c->has_clumps) return YA_E_STATE;
    pc_params P;
    P.bandWidth = c->P.bandWidth; P.GOCost = c->P.GOCost; P.GECost = c->P.GECost; P.RCost = c->P.RCost; P.MScore = c->P.MScore;
    P.minExtLength = c->P.minExtLength; P.maxROff = c->maxROff;
    uint32_t nj = 0;
    for (int s = 0; s < 2 * c->n_reads; s++) {
        const uint32_t nc = cb->clump_count[s], c0 = cb->clump_first[s];
        if (nc == 0 || nc == 0xFFFFFFFFu) continue;
        const uint32_t r = (uint32_t)(s >> 1);
        const int L = (int)(c->off[r + 1] - c->off[r]);
        const uint8_t *q = ((s & 1) ? c->rev : c->fwd) + c->off[r];
        for (uint32_t k = 0; k < nc; k++) {
            const ya_clump_rec rec = cb->clumps[c0 + k];
            memcpy(out->path + rec.first, cb->path + rec.first, rec.n * sizeof(ya_frag));
            ya_prep_rec pr;
            pr.gap_first = rec.first;
            pc_prepare_clump(&P, c->bases, q, L, r, s & 1, out->path + rec.first, (int)rec.n, out->gaps + rec.first, out->jobs, &nj,
                             (uint32_t)out->jobs_cap, &pr);
            out->prep[c0 + k] = pr;
        }
    }
    if (nj > out->jobs_cap) return YA_E_CAPACITY;
    out->n_jobs = nj;
    return 0;
}

int ya_sw_batch(ya_ctx *c, const ya_dp_job *jobs, int n, ya_dp_result *res, ya_op *ops, size_t ops_cap, size_t *ops_needed)
{
    size_t used = 0;
    int overflow = 0;
    c->pendingN = 0;
    uint64_t key = 0;
    if (memo_on()) {
        key = fnv(jobs, (size_t)n * sizeof(ya_dp_job), 14695981039346656037ull);
        memo *m = memo_find(key, 2);
        if (m && m->extra <= ops_cap) {
            memcpy(res, m->a, m->na); memcpy(ops, m->b, m->nb);
            if (ops_needed) *ops_needed = m->extra;
            return 0;
        }
    }
    ya_op *tmp = malloc(70000 * sizeof(ya_op));
    for (int i = 0; i < n; i++) {
        const ya_dp_job *j = &jobs[i];
        const uint8_t *codes = (j->strand ? c->rev : c->fwd) + c->off[j->read];
        int aq = 0, ar = 0, no = 0; int64_t cells = 0;
        int score = orc_dp(&c->P, c->bases, c->maxROff, codes, j->kind, j->rOff, j->rLen, j->qOff, j->qLen, &aq, &ar, tmp, 70000, &no, &cells);
        res[i].score = score; res[i].addedQLen = (uint16_t)aq; res[i].addedRLen = (uint16_t)ar; res[i].ops_off = (uint32_t)used; res[i].ops_n = (uint32_t)no;
        if (used + no <= ops_cap) memcpy(ops + used, tmp, no * sizeof(ya_op)); else overflow = 1;
        if (used + no > c->pendingCap) { c->pendingCap = 2 * (used + no) + 1024; c->pending = realloc(c->pending, c->pendingCap * sizeof(ya_op)); }
        memcpy(c->pending + used, tmp, no * sizeof(ya_op));
        used += no;
        c->ctr.dp_cells += cells; c->ctr.dp_jobs++;
    }
    free(tmp);
    if (ops_needed) *ops_needed = used;
    if (overflow) c->pendingN = used;
    if (memo_on()) {
        memo *m = memo_add(key, 2);
        m->na = (size_t)n * sizeof(ya_dp_result); m->a = dupmem(res, m->na);
        m->nb = used * sizeof(ya_op); m->b = dupmem(c->pending, m->nb);
        m->extra = used;
    }
    return overflow ? YA_E_CAPACITY : 0;
}

int ya_sw_fetch_ops(ya_ctx *c, ya_op *ops, size_t ops_cap)
{
    if (c->pendingN == 0) return YA_E_STATE;
    if (ops_cap < c->pendingN) return YA_E_CAPACITY;
    memcpy(ops, c->pending, c->pendingN * sizeof(ya_op));
    return 0;
}

int ya_perfect_ext(ya_ctx *c, const ya_dp_job *jobs, int n, uint16_t *count)
{
    for (int i = 0; i < n; i++) {
        const uint8_t *codes = (jobs[i].strand ? c->rev : c->fwd) + c->off[jobs[i].read];
        count[i] = (uint16_t)orc_perfect(c->bases, codes, jobs[i].rOff, jobs[i].qOff, jobs[i].qLen, jobs[i].kind == YA_DP_EXT_BWD ? -1 : 1);
    }
    return 0;
}
