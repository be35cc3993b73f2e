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
#include "../../yaha_b200/csrc/assemble_clumps.h"
#include "../../yaha_b200/csrc/finish_reads.h"
#include <math.h>

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
    /* ya_align_batch */
    ya_out_params out; int out_set;
    int n_seq; uint32_t *seq_start, *seq_len, *seq_name_off; char *seq_names;
    uint32_t *bpp_dist; int n_bpp, bpp_base;
    char *text; size_t text_pending;
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
ya_ctx *ya_open_peer(int device, const ya_ctx *src)
{
    ya_ctx *c = malloc(sizeof *c); *c = *src; c->fwd = c->rev = NULL; c->off = NULL; c->n_reads = 0; (void)device;
    c->pending = NULL; c->pendingN = c->pendingCap = 0;
    c->out_set = 0; c->seq_start = c->seq_len = c->seq_name_off = NULL; c->seq_names = NULL; c->bpp_dist = NULL; c->text = NULL; c->text_pending = 0;
    return c;
}
ya_ctx *ya_open_shared(const ya_ctx *src) { return ya_open_peer(0, src); }
int ya_peer_direct(const ya_ctx *c) { (void)c; return 0; }
int ya_bind_thread(const ya_ctx *c) { (void)c; return 0; }
int ya_set_priority(ya_ctx *c, int level) { (void)c; return level < 0 ? YA_E_ARG : 0; }
int ya_get_ext_intervals(ya_ctx *c, float *b, int cap, int *n) { (void)c; (void)b; (void)cap; *n = 0; return 0; }
ya_ctx *ya_open_build(int d, const ya_params *p, const uint8_t *b, size_t n, const uint32_t *s, const uint32_t *l, int ns, uint32_t mh, uint32_t sk)
{ (void)d; (void)p; (void)b; (void)n; (void)s; (void)l; (void)ns; (void)mh; (void)sk; return NULL; }
int ya_index_sizes(const ya_ctx *c, size_t *a, size_t *b) { *a = c->n_so; *b = c->n_roa; return 0; }
int ya_index_download(ya_ctx *c, uint32_t *so, uint32_t *roa) { (void)c; (void)so; (void)roa; return YA_E_STATE; }
void ya_close(ya_ctx *c)
{
    if (!c) return;
    free(c->fwd); free(c->rev); free(c->off); free(c->pending);
    free(c->seq_start); free(c->seq_len); free(c->seq_name_off); free(c->seq_names); free(c->bpp_dist); free(c->text);
    free(c);
}
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
        if (n > 1024) { out->clump_count[s] = 0xFFFFFFFFu; continue; }   /* as the device kernel: left to the host (clumps.cu) */
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
        if (used + no > ops_cap) overflow = 1; else if (no) memcpy(ops + used, tmp, no * sizeof(ya_op));
        if (used + no > c->pendingCap) { c->pendingCap = 2 * (used + no) + 1024; c->pending = realloc(c->pending, c->pendingCap * sizeof(ya_op)); }
        if (no) memcpy(c->pending + used, tmp, no * sizeof(ya_op));
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

/* ---- ya_align_batch on the CPU: the same headers the device kernels compile (form_clumps.h, prepare_clumps.h,
 * assemble_clumps.h, finish_reads.h) around the oracle's seed lookup and DP ---- */
int ya_set_output(ya_ctx *c, const ya_out_params *o, int n_seq, const char *const *names, const uint32_t *st, const uint32_t *ln)
{
    c->out = *o; c->out_set = 1; c->n_seq = n_seq;
    free(c->seq_start); free(c->seq_len); free(c->seq_name_off); free(c->seq_names); free(c->bpp_dist);
    c->seq_start = malloc(n_seq * 4); c->seq_len = malloc(n_seq * 4); c->seq_name_off = malloc((n_seq + 1) * 4);
    size_t tot = 0;
    for (int i = 0; i < n_seq; i++) tot += strlen(names[i]);
    c->seq_names = malloc(tot + 1);
    tot = 0;
    for (int i = 0; i < n_seq; i++) {
        c->seq_start[i] = st[i]; c->seq_len[i] = ln[i]; c->seq_name_off[i] = (uint32_t)tot;
        memcpy(c->seq_names + tot, names[i], strlen(names[i])); tot += strlen(names[i]);
    }
    c->seq_name_off[n_seq] = (uint32_t)tot;
    /* break-point penalty steps, as the CUDA library tabulates them (GraphPath.cpp:1018-1020) */
    int b0, b1;
    {
        double lg = log10(11.0); if (lg > o->maxBPLog) lg = o->maxBPLog; b0 = (int)(lg * o->BPCost + 0.5);
        lg = log10(4294967295.0); if (lg > o->maxBPLog) lg = o->maxBPLog; b1 = (int)(lg * o->BPCost + 0.5);
    }
    c->bpp_base = b0; c->n_bpp = 0; c->bpp_dist = malloc((size_t)(b1 > b0 ? b1 - b0 : 1) * 4);
    for (int b = b0 + 1; b <= b1; b++) {
        uint64_t lo = 11, hi = 0xFFFFFFFFull;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            double lg = log10((double)(uint32_t)mid); if (lg > o->maxBPLog) lg = o->maxBPLog;
            if ((int)(lg * o->BPCost + 0.5) >= b) hi = mid; else lo = mid + 1;
        }
        c->bpp_dist[c->n_bpp++] = (uint32_t)lo;
    }
    return 0;
}

int ya_align_fetch_text(ya_ctx *c, char *text, size_t cap)
{
    if (!c->text_pending) return YA_E_STATE;
    if (cap < c->text_pending) return YA_E_CAPACITY;
    memcpy(text, c->text, c->text_pending);
    return 0;
}

static int code_of_char(int ch)                                       /* Math.c:141-157 */
{
    static const char k[16] = {'T', 'C', 'A', 'G', 'N', 'B', 'D', 'H', 'K', 'M', 'R', 'S', 'V', 'W', 'X', 'Y'};
    if (ch == 'U' || ch == 'u') return 0;
    for (int i = 0; i < 16; i++) if (ch == k[i] || ch == k[i] + 32) return i;
    return 14;
}

int ya_align_batch(ya_ctx *c, ya_text_batch *b)
{
    const int n = b->n_reads;
    b->text_len = b->text_needed = 0; b->n_handed_back = 0; b->text_off[0] = 0;
    c->text_pending = 0;
    if (!c->out_set) return YA_E_STATE;
    if (n == 0) return 0;
    if (getenv("YA_MOCK_ALIGN_TOO_BIG")) return YA_E_STATE;       /* "does not fit one device pass": the host takes the call-by-call path */
    const uint64_t total = b->offsets[n];
    uint8_t *codes = malloc(total + 1);
    for (uint64_t i = 0; i < total; i++) codes[i] = (uint8_t)code_of_char((unsigned char)b->chars[i]);
    ya_read_batch rb; rb.n_reads = n; rb.codes = codes; rb.offsets = b->offsets;
    ya_reads_upload(c, &rb);
    free(codes);
    /* stages 1 + 2 */
    ya_frag_batch fb; memset(&fb, 0, sizeof fb);
    fb.frags_cap = (size_t)64 * n + 1024;
    fb.strands = malloc((size_t)2 * n * sizeof(ya_strand_frags));
    for (;;) {
        fb.frags = malloc(fb.frags_cap * sizeof(ya_frag)); fb.region = malloc(fb.frags_cap * 4);
        int rc = ya_seed_frags(c, &fb);
        if (rc == YA_E_CAPACITY) { free(fb.frags); free(fb.region); fb.frags_cap = fb.frags_needed + 1024; continue; }
        break;
    }
    const size_t nk = fb.n_frags, cap = nk + 16;
    ya_clump_batch cb; memset(&cb, 0, sizeof cb);
    cb.maxDesert = c->out.maxDesert; cb.minNonOverlap = c->out.minNonOverlap; cb.cap = cap;
    cb.clump_first = malloc((size_t)2 * n * 4); cb.clump_count = malloc((size_t)2 * n * 4);
    cb.clumps = malloc(cap * sizeof(ya_clump_rec)); cb.path = malloc(cap * sizeof(ya_frag));
    ya_form_clumps(c, &cb);
    ya_prep_batch pb; memset(&pb, 0, sizeof pb);
    pb.cap = cap; pb.jobs_cap = 3 * cap + 16;
    pb.prep = malloc(cap * sizeof(ya_prep_rec)); pb.gaps = malloc(cap * sizeof(ya_gap_rec)); pb.path = malloc(cap * sizeof(ya_frag));
    pb.jobs = malloc(pb.jobs_cap * sizeof(ya_dp_job));
    ya_prepare_clumps(c, &pb);
    /* the first DP round */
    ya_dp_result *res = malloc((pb.n_jobs + 1) * sizeof(ya_dp_result));
    size_t opsCap = 64 * pb.n_jobs + 1024, need = 0;
    ya_op *rops = malloc(opsCap * sizeof(ya_op));
    if (ya_sw_batch(c, pb.jobs, (int)pb.n_jobs, res, rops, opsCap, &need) == YA_E_CAPACITY) {
        opsCap = need + 16; rops = realloc(rops, opsCap * sizeof(ya_op));
        ya_sw_fetch_ops(c, rops, opsCap);
    }
    /* splice + score every clump */
    ac_params AP;
    AP.GOCost = c->P.GOCost; AP.GECost = c->P.GECost; AP.RCost = c->P.RCost; AP.MScore = c->P.MScore; AP.minExtLength = c->P.minExtLength;
    AP.minRawScore = c->out.minRawScore; AP.maxROff = c->maxROff; AP.minIdentity = c->out.minIdentity;
    ya_asm_rec *recs = calloc(cap, sizeof(ya_asm_rec));
    size_t asmCap = need + 2 * nk + 64, asmUsed = 0;
    ya_op *asmOps = malloc(asmCap * sizeof(ya_op));
    for (int s = 0; s < 2 * n; s++) {
        const uint32_t nc = cb.clump_count[s], c0 = cb.clump_first[s];
        if (nc == 0 || nc == 0xFFFFFFFFu) continue;
        const int r = s >> 1;
        const int L = (int)(c->off[r + 1] - c->off[r]);
        const uint8_t *q = ((s & 1) ? c->rev : c->fwd) + c->off[r];
        for (uint32_t k = 0; k < nc; k++) {
            const ya_clump_rec cr = cb.clumps[c0 + k];
            const ya_prep_rec *pr = &pb.prep[c0 + k];
            const ya_gap_rec *g = pb.gaps + pr->gap_first;
            const uint32_t bound = ac_ops_bound((int)cr.n, g, pr->n_gaps, pr, res, rops);
            if (asmUsed + bound > asmCap) { fprintf(stderr, "mock: run array too small\n"); abort(); }
            if (ac_assemble_clump(&AP, c->bases, q, L, pb.path + cr.first, (int)cr.n, g, pr->n_gaps, pr, res, rops, asmOps + asmUsed, &recs[c0 + k]) != 0) {
                fprintf(stderr, "mock: extension plan diverged\n"); abort();
            }
            recs[c0 + k].ops_off = (uint32_t)asmUsed;
            asmUsed += bound;
        }
    }
    if (getenv("YA_MOCK_TRACE")) fprintf(stderr, "mock: assembled, %zu runs\n", asmUsed);
    /* finish every read */
    fr_params F; memset(&F, 0, sizeof F);
    F.GOCost = c->P.GOCost; F.GECost = c->P.GECost; F.RCost = c->P.RCost; F.MScore = c->P.MScore;
    F.OQC = c->out.OQC; F.FBS = c->out.FBS; F.OQCMinNonOverlap = c->out.OQCMinNonOverlap; F.BPCost = c->out.BPCost; F.maxBPLog = c->out.maxBPLog;
    F.FBS_PSLength = c->out.FBS_PSLength; F.FBS_PSScore = c->out.FBS_PSScore; F.hardClip = c->out.hardClip; F.fastq = c->out.fastq;
    F.n_seq = c->n_seq; F.seq_start = c->seq_start; F.seq_len = c->seq_len; F.seq_name_off = c->seq_name_off; F.seq_names = c->seq_names;
    F.bpp_base = c->bpp_base; F.n_bpp = c->n_bpp; F.bpp_dist = c->bpp_dist;
    size_t textCap = 1 << 20, textLen = 0;
    free(c->text); c->text = malloc(textCap);
    for (int r = 0; r < n; r++) {
        const uint64_t base = c->off[r];
        const int L = (int)(c->off[r + 1] - base);
        fr_clump cl[FR_MAX_NODES];
        int m = 0, hand = 0;
        for (int st = 0; st < 2 && !hand; st++) {
            const int s = 2 * r + st;
            const uint32_t nc = cb.clump_count[s], c0 = cb.clump_first[s];
            if (nc == 0xFFFFFFFFu) { hand = 1; break; }
            for (uint32_t k = 0; k < nc; k++) {
                const ya_asm_rec *a = &recs[c0 + k];
                if (a->verdict == YA_ASM_SPLIT) { hand = 1; break; }
                if (a->verdict != YA_ASM_SCORED) continue;
                if (m == FR_MAX_NODES) { hand = 1; break; }
                cl[m].rec = a; cl[m].ops = asmOps + a->ops_off; cl[m].reversed = st; m++;
            }
        }
        b->text_off[r] = textLen;
        if (getenv("YA_MOCK_TRACE")) fprintf(stderr, "mock: read %d clumps %d hand %d\n", r, m, hand);
        if (!hand && m > 0) {
            fr_node g[FR_MAX_NODES]; fr_out o[FR_MAX_NODES];
            int prim = 0;
            const int k = fr_finish_read(&F, c->fwd + base, L, cl, m, g, o, &prim);
            if (k < 0) hand = 1;
            for (int q = 0; q < k; q++) {
                const size_t len = fr_format_record(&F, c->bases, b->ids + b->id_off[r], (int)(b->id_off[r + 1] - b->id_off[r]), b->chars + base,
                                                    b->quals ? b->quals + base : NULL, c->rev + base, L, &cl[o[q].clump], &o[q], prim, NULL);
                if (textLen + len > textCap) { textCap = 2 * (textLen + len); c->text = realloc(c->text, textCap); }
                const size_t wrote = fr_format_record(&F, c->bases, b->ids + b->id_off[r], (int)(b->id_off[r + 1] - b->id_off[r]), b->chars + base,
                                                      b->quals ? b->quals + base : NULL, c->rev + base, L, &cl[o[q].clump], &o[q], prim, c->text + textLen);
                if (wrote != len) { fprintf(stderr, "mock: record length changed between the counting and the writing pass\n"); abort(); }
                textLen += len;
            }
        }
        b->status[r] = (uint8_t)hand;
        b->n_handed_back += hand;
    }
    b->text_off[n] = textLen;
    b->text_len = b->text_needed = textLen;
    free(fb.strands); free(fb.frags); free(fb.region); free(cb.clump_first); free(cb.clump_count); free(cb.clumps); free(cb.path);
    free(pb.prep); free(pb.gaps); free(pb.path); free(pb.jobs); free(res); free(rops); free(recs); free(asmOps);
    c->has_seed = 0; c->has_clumps = 0;
    if (textLen > b->text_cap) { c->text_pending = textLen; return YA_E_CAPACITY; }
    memcpy(b->text, c->text, textLen);
    return 0;
}
