// ctx.cu -- context life-cycle, index residency and read-batch upload for libyaha_b200.so.
//
// Replaces the per-run set-up of processQueryFile (Query.c:551-640: mmap of .nib2 and index,
// pointer carving) with HBM-resident copies, and the per-read buffer fill of readNextQuery
// (Query.c:161-168: forward + reverse-complement code buffers) with one batched upload plus
// a device kernel that derives the reverse-complement strand.
#include "common.cuh"
#include <condition_variable>
#include <mutex>
#include <sched.h>
#include <sys/prctl.h>
#include <time.h>
#include <algorithm>

static thread_local std::string g_open_err;

// complement of a 4-bit code (Math.c:155: fourBitCompCodes)
__constant__ uint8_t c_comp[16] = {2, 3, 0, 1, 4, 12, 7, 6, 9, 8, 15, 11, 5, 13, 14, 10};

// Reverse-complement strand of every read (Query.c:164-167): one warp per read, both accesses coalesced.
__global__ void revcomp_kernel(const uint8_t *__restrict__ fwd, uint8_t *__restrict__ rev,
                               const uint64_t *__restrict__ off, int n_reads, uint64_t total)
{
    const int lane = threadIdx.x & 31;
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    (void)total;
    for (int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n_reads; r += warps) {
        const uint64_t s = off[r], e = off[r + 1];
        for (uint64_t i = s + lane; i < e; i += 32) rev[s + (e - 1 - i)] = c_comp[fwd[i] & 15];
    }
}

static int check_params(const ya_params *p, std::string &why)
{
    if (p->wordLen < 1 || p->wordLen > 16) { why = "wordLen must be in 1..16"; return YA_E_ARG; }
    if (p->bandWidth < 0 || p->bandWidth > 4000) { why = "bandWidth out of range"; return YA_E_ARG; }
    if (p->maxGap < 0 || p->maxGap > 16383 || p->maxIntron < 0 || p->maxIntron > 16383) {
        why = "maxGap/maxIntron above 16383 are not supported (14-bit traceback run lengths)";
        return YA_E_ARG;
    }
    if (p->maxHits < 0 || p->maxHits > 65525) { why = "maxHits out of range"; return YA_E_ARG; }
    return YA_OK;
}

static ya_ctx *open_common(int device, const ya_params *params)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_open_err = "no CUDA device available: yaha_b200 has no CPU fallback";
        return nullptr;
    }
    if (device < 0 || device >= ndev) { g_open_err = "bad device ordinal"; return nullptr; }
    std::string why;
    if (!params || check_params(params, why) != YA_OK) { g_open_err = "bad ya_params: " + why; return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { g_open_err = "cudaGetDeviceProperties failed"; return nullptr; }
    if (prop.major < 10) {
        g_open_err = std::string("device '") + prop.name + "' is not sm_100 class; this library ships sm_100a code only";
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_open_err = "cudaSetDevice failed"; return nullptr; }
    {
        // YA_SYNC=block: the thread that drives a context sleeps in its stream synchronisations (see ya_stream_wait)
        const char *e = getenv("YA_SYNC");
        if (e && strcmp(e, "block") == 0) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
    }
    {
        // Every hot access of this library is a random gather (starting-offset table, ROA lists,
        // back-pointer walk): ask L2 to fetch 32 B sectors instead of whole 128 B lines so that a
        // probe costs one DRAM sector, not four (measured: 4.85 sectors/probe at the default).
        const char *e = getenv("YA_L2_FETCH");
        size_t gran = e ? (size_t)atoi(e) : 32;
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    }
    {
        // freed scratch stays in the device's stream-ordered pool (never handed back to the driver mid-run)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    ya_ctx *c = new ya_ctx();
    c->device = device;
    c->P = *params;
    // Two streams: latency-critical work (read upload, seed stage, DP calls with few jobs) runs at the highest
    // priority so that its small kernels are not queued behind another pipeline's bulk DP launches.
    int prLeast = 0, prGreatest = 0;
    cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest);
    if (cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, prGreatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->bulk_stream, cudaStreamNonBlocking, prLeast) != cudaSuccess) {
        g_open_err = "cudaStreamCreate failed"; delete c; return nullptr;
    }
    c->stream = c->own_stream;
    c->prio_streams[0] = c->own_stream;
    for (int i = 0; i < 6; i++) cudaEventCreate(&c->ev[i]);
    return c;
}

// The over-read quirk of QueryMatch.c:62-67 can only trigger for a k-mer whose LAST occurrence lies
// below the query offset, i.e. inside the first 32 767 reference bases.  A 2^20-bit filter of the
// k-mers that occur there lets the seed kernel skip the extra ROA gather for everything else.
#define YA_LOWMASK_WORDS (1u << 15)
int ya_build_lowmask(ya_ctx *c, const uint8_t *bases, size_t n_base_bytes)
{
    std::vector<uint32_t> m(YA_LOWMASK_WORDS, 0u);
    const int K = c->P.wordLen;
    const uint32_t mask = 0xFFFFFFFFu >> (32 - 2 * K);
    const uint64_t limit = std::min<uint64_t>((uint64_t)n_base_bytes * 2, (uint64_t)32768 + K);
    uint32_t h = 0; int good = 0;
    for (uint64_t pos = 0; pos < limit; pos++) {
        const uint8_t b = bases[pos >> 1];
        const uint32_t code = (pos & 1) ? (b & 15u) : (b >> 4);
        if (code > 3) { good = 0; h = 0; continue; }
        h = ((h << 2) | code) & mask;
        if (++good >= K) m[(h & 0xFFFFFu) >> 5] |= 1u << (h & 31u);
    }
    if (!c->d_lowmask && cudaMalloc(&c->d_lowmask, YA_LOWMASK_WORDS * 4) != cudaSuccess) return YA_E_CUDA;
    if (cudaMemcpy(c->d_lowmask, m.data(), YA_LOWMASK_WORDS * 4, cudaMemcpyHostToDevice) != cudaSuccess) return YA_E_CUDA;
    return YA_OK;
}

ya_ctx *ya_open_common_for_index(int device, const ya_params *params) { return open_common(device, params); }
void ya_set_open_error(const std::string &m) { g_open_err = m; }

extern "C" ya_ctx *ya_open(int device, const ya_params *params,
                           const uint32_t *so, size_t n_so, const uint32_t *roa, size_t n_roa,
                           const uint8_t *bases, size_t n_base_bytes, uint32_t maxROff)
{
    ya_ctx *c = open_common(device, params);
    if (!c) return nullptr;
    size_t want_so = ((size_t)1 << (2 * params->wordLen)) + 1;
    if (!so || n_so != want_so) {
        g_open_err = "starting-offset table must have 4^wordLen + 1 entries"; ya_close(c); return nullptr;
    }
    cudaError_t e;
    // +8 words of zero padding after the ROA: the over-read of QueryMatch.c:62-67 is bounded
    // by n_roa in our kernels, the padding only keeps vector loads in bounds.
    if ((e = cudaMalloc(&c->d_so, n_so * 4)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_roa, (n_roa + 8) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_bases, n_base_bytes + 64)) != cudaSuccess) {
        g_open_err = std::string("cudaMalloc(index): ") + cudaGetErrorString(e); ya_close(c); return nullptr;
    }
    if ((e = cudaMemcpy(c->d_so, so, n_so * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (n_roa && (e = cudaMemcpy(c->d_roa, roa, n_roa * 4, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (e = cudaMemset(c->d_roa + n_roa, 0, 32)) != cudaSuccess ||
        (e = cudaMemcpy(c->d_bases, bases, n_base_bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemset(c->d_bases + n_base_bytes, 0xEE, 64)) != cudaSuccess) {
        g_open_err = std::string("cudaMemcpy(index): ") + cudaGetErrorString(e); ya_close(c); return nullptr;
    }
    c->n_so = n_so; c->n_roa = n_roa; c->n_base_bytes = n_base_bytes; c->maxROff = maxROff;
    if (ya_build_lowmask(c, bases, n_base_bytes) != YA_OK) { g_open_err = "cudaMalloc(lowmask) failed"; ya_close(c); return nullptr; }
    return c;
}

extern "C" ya_ctx *ya_open_peer(int device, const ya_ctx *src)
{
    if (!src) { g_open_err = "ya_open_peer: null source"; return nullptr; }
    ya_ctx *c = open_common(device, &src->P);
    if (!c) return nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_so, src->n_so * 4)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_roa, (src->n_roa + 8) * 4)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_bases, src->n_base_bytes + 64)) != cudaSuccess) {
        g_open_err = std::string("cudaMalloc(index): ") + cudaGetErrorString(e); ya_close(c); return nullptr;
    }
    // Device-to-device over NVLink: with peer access enabled in both directions cudaMemcpyPeer is a direct copy between the
    // two HBMs through NVSwitch; without it (not supported, or refused) the driver stages the copy through host memory.
    {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
            cudaSetDevice(device);
            cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            cudaGetLastError();
            cudaSetDevice(src->device);
            pe = cudaDeviceEnablePeerAccess(device, 0);
            cudaGetLastError();
            cudaSetDevice(device);
        }
        c->peer_direct = can != 0;
    }
    if ((e = cudaMemcpyPeer(c->d_so, device, src->d_so, src->device, src->n_so * 4)) != cudaSuccess ||
        (e = cudaMemcpyPeer(c->d_roa, device, src->d_roa, src->device, (src->n_roa + 8) * 4)) != cudaSuccess ||
        (e = cudaMemcpyPeer(c->d_bases, device, src->d_bases, src->device, src->n_base_bytes + 64)) != cudaSuccess) {
        g_open_err = std::string("cudaMemcpyPeer(index): ") + cudaGetErrorString(e); ya_close(c); return nullptr;
    }
    c->n_so = src->n_so; c->n_roa = src->n_roa; c->n_base_bytes = src->n_base_bytes; c->maxROff = src->maxROff;
    if (cudaMalloc(&c->d_lowmask, YA_LOWMASK_WORDS * 4) != cudaSuccess ||
        cudaMemcpyPeer(c->d_lowmask, device, src->d_lowmask, src->device, YA_LOWMASK_WORDS * 4) != cudaSuccess) {
        g_open_err = "lowmask peer copy failed"; ya_close(c); return nullptr;
    }
    return c;
}

cudaEvent_t ya_device_base_event(int device, cudaStream_t st)
{
    static cudaEvent_t base[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> g(mu);
    cudaEvent_t &e = base[(unsigned)device & 63u];
    if (!e) { cudaEventCreate(&e); cudaEventRecord(e, st); cudaEventSynchronize(e); }
    return e;
}

// the span of a bulk extension launch on the device's common time axis (the pipelines of a device overlap their launches: the
// UNION of the spans is the time the kernel class occupies the device, which per-launch durations cannot tell)
void ya_note_ext_interval(ya_ctx *c, cudaEvent_t a, cudaEvent_t b)
{
    cudaEvent_t base = ya_device_base_event(c->device, c->stream);
    float t0 = 0, t1 = 0;
    if (cudaEventElapsedTime(&t0, base, a) == cudaSuccess && cudaEventElapsedTime(&t1, base, b) == cudaSuccess && c->ext_iv.size() < (1u << 20)) {
        c->ext_iv.push_back(t0); c->ext_iv.push_back(t1);
    } else cudaGetLastError();
}

extern "C" int ya_get_ext_intervals(ya_ctx *c, float *start_end_ms, int cap_pairs, int *n_pairs)
{
    if (!c || !n_pairs) return YA_E_ARG;
    const int n = (int)(c->ext_iv.size() / 2);
    *n_pairs = n;
    if (n > cap_pairs || (n && !start_end_ms)) return YA_E_CAPACITY;
    if (n) memcpy(start_end_ms, c->ext_iv.data(), (size_t)n * 2 * sizeof(float));
    c->ext_iv.clear();
    return YA_OK;
}

extern "C" int ya_peer_direct(const ya_ctx *c) { return c && c->peer_direct ? 1 : 0; }

extern "C" int ya_bind_thread(const ya_ctx *c)
{
    if (!c) return YA_E_ARG;
    return cudaSetDevice(c->device) == cudaSuccess ? YA_OK : YA_E_CUDA;
}

extern "C" void *ya_host_alloc(size_t bytes)
{
    void *p = nullptr;
    const double t0 = ya_now();
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ya_note_alloc("host_alloc", 0, bytes, t0);
    return p;
}
extern "C" void ya_host_free(void *p) { if (p) { const double t0 = ya_now(); cudaFreeHost(p); ya_note_alloc("host_free", 0, 0, t0); } }

thread_local cudaStream_t ya_tls_alloc_stream = nullptr;
double ya_now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
void ya_note_alloc(const char *kind, size_t old_bytes, size_t new_bytes, double t0)
{
    static const bool on = getenv("YA_ALLOC_LOG") != nullptr;
    if (on) fprintf(stderr, "ya_alloc %s %zu -> %zu bytes, %.3f ms, at %.6f\n", kind, old_bytes, new_bytes, (ya_now() - t0) * 1e3, ya_now());
}

static int ya_sync_mode()
{
    // 0 nap (default): poll, sleeping ~20 us between polls (timer slack lowered for this thread)
    // 1 the driver's own synchronisation ("spin", or "block" with cudaDeviceScheduleBlockingSync)
    // 2 yield: poll and sched_yield between polls
    static const int mode = [] {
        const char *e = getenv("YA_SYNC");
        if (!e) return 0;
        if (strcmp(e, "spin") == 0 || strcmp(e, "block") == 0) return 1;
        if (strcmp(e, "yield") == 0) return 2;
        return 0;
    }();
    return mode;
}
template <class Query> static cudaError_t ya_poll(Query q)
{
    static const long nap_ns = [] { const char *e = getenv("YA_NAP_US"); return (long)(e ? atoi(e) : 20) * 1000L; }();
    const int mode = ya_sync_mode();
    static thread_local bool slackSet = false;
    if (mode == 0 && !slackSet) { prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0); slackSet = true; }
    for (int polls = 0;; polls++) {
        cudaError_t e = q();
        if (e != cudaErrorNotReady) return e;
        if (mode == 2 || polls < 4) sched_yield();
        else { struct timespec ts = {0, nap_ns}; nanosleep(&ts, nullptr); }
    }
}
cudaError_t ya_stream_wait(cudaStream_t st)
{
    if (ya_sync_mode() == 1) return cudaStreamSynchronize(st);
    return ya_poll([st] { return cudaStreamQuery(st); });
}
cudaError_t ya_event_wait(cudaEvent_t ev)
{
    if (ya_sync_mode() == 1) return cudaEventSynchronize(ev);
    return ya_poll([ev] { return cudaEventQuery(ev); });
}
std::mutex &ya_bulk_mutex(int device)
{
    static std::mutex mu[64];
    return mu[(unsigned)device & 63u];
}

std::mutex &ya_device_mutex(int device)
{
    static std::mutex mu[64];
    return mu[(unsigned)device & 63u];
}
bool ya_device_turns()
{
    static const bool on = getenv("YA_GPU_LOCK") != nullptr;      // off by default: contexts of one device overlap their phases
    return on;
}

extern "C" ya_ctx *ya_open_shared(const ya_ctx *src)
{
    if (!src) { g_open_err = "ya_open_shared: null source"; return nullptr; }
    ya_ctx *c = open_common(src->device, &src->P);
    if (!c) return nullptr;
    c->d_so = src->d_so; c->d_roa = src->d_roa; c->d_bases = src->d_bases; c->d_lowmask = src->d_lowmask;
    c->n_so = src->n_so; c->n_roa = src->n_roa; c->n_base_bytes = src->n_base_bytes; c->maxROff = src->maxROff;
    c->owns_index = false;
    return c;
}

extern "C" void ya_close(ya_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->owns_index) {
        if (c->d_so) cudaFree(c->d_so);
        if (c->d_roa) cudaFree(c->d_roa);
        if (c->d_bases) cudaFree(c->d_bases);
        if (c->d_lowmask) cudaFree(c->d_lowmask);
    }
    DevBuf *bufs[] = {&c->d_codes_fwd, &c->d_codes_rev, &c->d_read_off, &c->d_seg_probe_off, &c->d_cnt, &c->d_soff,
                      &c->d_hit_off, &c->d_keys0, &c->d_keys1, &c->d_scan_tmp, &c->d_hist, &c->d_fragflag, &c->d_fragidx,
                      &c->d_frags_all, &c->d_frag_seg, &c->d_regflag, &c->d_regidx, &c->d_regstart, &c->d_keep,
                      &c->d_keepidx, &c->d_frags_out, &c->d_region_out, &c->d_strand_out, &c->d_misc, &c->d_jobs,
                      &c->d_jobout, &c->d_tb, &c->d_rows, &c->d_ops_raw, &c->d_ops_cnt, &c->d_ops_off, &c->d_ops_out, &c->d_res,
                      &c->d_fc_count, &c->d_fc_work, &c->d_fc_tmp, &c->d_fc_path, &c->d_fc_nodes, &c->d_fc_used, &c->d_fc_clumps, &c->d_fc_slot,
                      &c->d_pc_path, &c->d_pc_gaps, &c->d_pc_prep, &c->d_pc_jobs, &c->d_dpr, &c->d_chars, &c->d_quals, &c->d_ids, &c->d_fin,
                      &c->d_asm_recs, &c->d_asm_ops, &c->d_fr_outs, &c->d_text, &c->d_out_tab};
    for (DevBuf *b : bufs) b->release();
    PinBuf *pins[] = {&c->h_stage, &c->h_stage2, &c->h_stage3, &c->h_jobs, &c->h_res, &c->h_ops};
    for (PinBuf *b : pins) b->release();
    for (int i = 0; i < 6; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (cudaStream_t ps : c->prio_streams) if (ps) cudaStreamDestroy(ps);
    if (c->bulk_stream) cudaStreamDestroy(c->bulk_stream);
    delete c;
}

extern "C" const char *ya_last_error(const ya_ctx *c)
{
    return c ? c->err.c_str() : g_open_err.c_str();
}

extern "C" int ya_set_params(ya_ctx *c, const ya_params *p)
{
    if (!c || !p) return YA_E_ARG;
    std::string why;
    if (check_params(p, why) != YA_OK) return ya_fail(c, YA_E_ARG, why);
    if (p->wordLen != c->P.wordLen) return ya_fail(c, YA_E_ARG, "wordLen is fixed by the resident index");
    c->P = *p;
    return YA_OK;
}

extern "C" int ya_set_stream(ya_ctx *c, void *s)
{
    if (!c) return YA_E_ARG;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return YA_OK;
}

// Urgency of the context's own work among the contexts sharing its device: level 0 (default) is the device's greatest stream
// priority, every further level one step lower (clamped one above the bulk stream's).  The host program gives the batches in
// flight on one device descending urgency in input order, so that they FINISH one after the other -- the ordered writer then
// works on batch k while batch k+1 still computes -- instead of all at once.  To be called between batches (context idle).
extern "C" int ya_set_priority(ya_ctx *c, int level)
{
    if (!c || level < 0) return YA_E_ARG;
    if (level > 3) level = 3;
    if (level == c->prio_level) return YA_OK;
    YA_CUDA(c, cudaSetDevice(c->device));
    if (!c->prio_streams[level]) {
        int prLeast = 0, prGreatest = 0;
        YA_CUDA(c, cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
        int pr = prGreatest + level;                                  // (numerically greater = less urgent)
        if (pr > prLeast - 1) pr = prLeast - 1;
        if (pr < prGreatest) pr = prGreatest;
        YA_CUDA(c, cudaStreamCreateWithPriority(&c->prio_streams[level], cudaStreamNonBlocking, pr));
    }
    YA_CUDA(c, cudaStreamSynchronize(c->own_stream));                 // (idle by contract; scratch freed there is reusable after this)
    const bool installed = (c->stream == c->own_stream);
    c->own_stream = c->prio_streams[level];
    c->prio_level = level;
    if (installed) c->stream = c->own_stream;
    return YA_OK;
}

extern "C" int ya_reads_upload(ya_ctx *c, const ya_read_batch *b)
{
    if (!c || !b || b->n_reads < 0 || (b->n_reads && (!b->codes || !b->offsets))) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    if (b->n_reads > (1 << 21)) return ya_fail(c, YA_E_ARG, "at most 2^21 reads per batch");
    AllocScope allocScope(c->stream);
    c->n_reads = b->n_reads;
    c->h_read_off.assign(b->offsets, b->offsets + b->n_reads + 1);
    if (c->h_read_off[0] != 0) return ya_fail(c, YA_E_ARG, "offsets[0] must be 0");
    for (int r = 0; r < b->n_reads; r++) {
        uint64_t L = c->h_read_off[r + 1] - c->h_read_off[r];
        if (c->h_read_off[r + 1] < c->h_read_off[r] || L > 32767)
            return ya_fail(c, YA_E_ARG, "read length must be in 0..32767 (16-bit query offsets, Math.h:104)");
    }
    uint64_t total = c->h_read_off[b->n_reads];
    // query codes are addressed with 32-bit indices downstream (DevJob.qIdx, the probe offsets of ya_seed_frags)
    if (total >= 0xFFFF0000ull) { c->n_reads = 0; return ya_fail(c, YA_E_ARG, "at most 2^32 - 65536 bases per batch (32-bit code offsets): split the batch"); }
    c->total_bases = total;
    YA_CUDA(c, c->d_codes_fwd.reserve(total + 64));
    YA_CUDA(c, c->d_codes_rev.reserve(total + 64));
    YA_CUDA(c, c->d_read_off.reserve((size_t)(b->n_reads + 1) * 8));
    YA_CUDA(c, cudaMemcpyAsync(c->d_read_off.p, b->offsets, (size_t)(b->n_reads + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (total) {
        YA_CUDA(c, cudaMemcpyAsync(c->d_codes_fwd.p, b->codes, total, cudaMemcpyHostToDevice, c->stream));
        int blocks = (b->n_reads + 7) / 8; if (blocks > 148 * 16) blocks = 148 * 16;
        revcomp_kernel<<<blocks, 256, 0, c->stream>>>(c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(),
                                                     c->d_read_off.as<uint64_t>(), b->n_reads, total);
        c->ctr.launches++;
        YA_CUDA(c, cudaGetLastError());
    }
    YA_CUDA(c, ya_stream_wait(c->stream));
    return YA_OK;
}

extern "C" int ya_get_counters(ya_ctx *c, ya_counters *out)
{
    if (!c || !out) return YA_E_ARG;
    *out = c->ctr;
    c->ctr = ya_counters{};
    return YA_OK;
}
