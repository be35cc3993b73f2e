// index.cu -- build yaha's k-mer index on the device, in the reference's exact layout.
//
// Replaces indexFile (Index.c:49-331) for every -L / -S / -H:
//   pass 1+2 (count, fill; Index.c:95-242)  one thread per reference position hashes its window; positions the
//            reference's walk would not visit (window holds a non-ACGT code, or the offset is off the -S grid) are
//            skipped; histogram into the starting-offset table -> exclusive scan (Index.c:185-194) -> stable radix sort
//            of (hash, position) keys by hash, which leaves each k-mer's list in ascending offset order exactly like
//            the reference's second pass.
//   the -S grid (Index.c:105-128): the walk starts at the sequence's first base and advances by skipDist; after a
//            window with a bad code it resumes at the first multiple of skipDist (of the GLOBAL offset) behind that
//            run of bad codes.  So a clean window at p is indexed iff (p - seqStart) % S == 0 while no bad code of
//            the sequence lies below p, and iff p % S == 0 afterwards.
//   pass 3  (sample; Index.c:271-315, Math.c:274-343)  a k-mer with more than maxHits occurrences keeps maxHits of
//            them, chosen by Floyd's algorithm from ONE xorshift stream that runs through the over-full k-mers in hash
//            order.  The stream is sequential by construction, so the draws are made on the host for the (few) over-full
//            lists -- the device finds them, the host returns one rank-or-dropped word per occurrence of those lists,
//            and the device rewrites the offset table and compacts the ROA.
#include "common.cuh"
#include <algorithm>

__global__ void index_firstbad_kernel(const uint8_t *__restrict__ bases, uint32_t seq_start, uint32_t seq_len, uint32_t *__restrict__ firstbad)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0xFFFFFFFFu;
    if (t < seq_len) {
        const uint32_t o = seq_start + t;
        const uint32_t b = bases[o >> 1];
        if (((o & 1) ? (b & 15u) : (b >> 4)) > 3u) m = o;
    }
    for (int d = 16; d; d >>= 1) { const uint32_t x = __shfl_xor_sync(0xffffffffu, m, d); if (x < m) m = x; }
    if ((threadIdx.x & 31) == 0 && m != 0xFFFFFFFFu) atomicMin(firstbad, m);
}

__global__ void index_hash_kernel(const uint8_t *__restrict__ bases, uint32_t seq_start, uint32_t n_pos, int K, uint32_t skip,
                                  const uint32_t *__restrict__ firstbad, uint32_t *__restrict__ counts,
                                  uint64_t *__restrict__ keys, uint32_t key_base)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pos) return;
    uint32_t pos = seq_start + t;
    uint32_t h = 0, bad = 0;
    for (int k = 0; k < K; k++) {
        uint32_t o = pos + k;
        uint32_t b = bases[o >> 1];
        uint32_t c = (o & 1) ? (b & 15u) : (b >> 4);
        bad |= c;
        h = (h << 2) | (c & 3u);
    }
    bool take = bad < 4;
    if (take && skip > 1) take = (pos > *firstbad) ? (pos % skip == 0) : (t % skip == 0);       // Index.c:105-128
    if (take) {
        atomicAdd(&counts[h], 1u);
        keys[key_base + t] = ((uint64_t)h << 32) | pos;
    } else {
        keys[key_base + t] = ~0ull;                      // sorts behind every real k-mer
    }
}

// The sort orders keys by their hash bits only, and a skipped position's key (all ones) carries the largest hash there is:
// behind the sort the keys of the all-G k-mer and the skipped ones are interleaved (in position order).  These two kernels
// move the real ones in front, keeping their order.
__global__ void index_tail_flag_kernel(const uint64_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ flag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = keys[i] != ~0ull;
}
__global__ void index_tail_scatter_kernel(const uint64_t *__restrict__ keys, uint32_t n, const uint32_t *__restrict__ flag,
                                          const uint32_t *__restrict__ idx, uint64_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[idx[i]] = keys[i];
}

__global__ void index_take_pos_kernel(const uint64_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ roa)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) roa[i] = (uint32_t)keys[i];
}

// k-mers with more than maxHits occurrences: (hash, count) pairs, in no particular order (the host sorts the few there are)
struct OverFull { uint32_t hash, count, start; };
__global__ void index_overfull_kernel(const uint32_t *__restrict__ so, size_t n_kmers, uint32_t maxHits,
                                      OverFull *__restrict__ list, uint32_t cap, uint32_t *__restrict__ n_list)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n_kmers; i += step) {
        const uint32_t c = so[i + 1] - so[i];
        if (c > maxHits) {
            const uint32_t k = atomicAdd(n_list, 1u);
            if (k < cap) { list[k].hash = (uint32_t)i; list[k].count = c; list[k].start = so[i]; }
        }
    }
}

// After sampling: per over-full k-mer (ascending hash) its first word in `rank` and the number of occurrences dropped
// from all over-full k-mers before it.
struct OverPlan { uint32_t hash, rank_off, dropped_before, dropped; };

__device__ __forceinline__ int over_find(const OverPlan *__restrict__ plan, int n, uint32_t h)   // last entry with hash <= h, or -1
{
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (plan[mid].hash <= h) lo = mid + 1; else hi = mid; }
    return lo - 1;
}

// new offset table: every list starts `dropped before it` entries earlier (Index.c:300-301)
__global__ void index_resample_so_kernel(uint32_t *__restrict__ so, size_t n_so, const OverPlan *__restrict__ plan, int n_plan)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n_so; i += step) {
        const int k = over_find(plan, n_plan, i > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)i);
        if (k < 0) continue;
        uint32_t d = plan[k].dropped_before;
        if (i > plan[k].hash || i == n_so - 1) d += plan[k].dropped;     // (entry 4^K is the total; hashes are < 4^K)
        so[i] -= d;
    }
}

// sorted (hash, position) keys -> sampled ROA.  so_old is the offset table before sampling (the key's index in its list).
__global__ void index_resample_roa_kernel(const uint64_t *__restrict__ keys, uint32_t n, const uint32_t *__restrict__ so_new,
                                          const OverPlan *__restrict__ plan, int n_plan, const uint32_t *__restrict__ so_old_of_plan,
                                          const uint32_t *__restrict__ rank, uint32_t *__restrict__ roa)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t key = keys[j];
    const uint32_t h = (uint32_t)(key >> 32);
    const int k = over_find(plan, n_plan, h);
    if (k < 0) { roa[j] = (uint32_t)key; return; }
    if (plan[k].hash == h) {
        const uint32_t r = rank[plan[k].rank_off + (j - so_old_of_plan[k])];
        if (r != 0xFFFFFFFFu) roa[so_new[h] + r] = (uint32_t)key;
    } else roa[j - plan[k].dropped_before - plan[k].dropped] = (uint32_t)key;
}

// Marsaglia xorshift, Math.c:274-284
static inline uint32_t rand_bits(uint32_t *s)
{
    const uint32_t t = s[0] ^ (s[0] >> 7);
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3]; s[3] = s[4];
    s[4] = (s[4] ^ (s[4] << 6)) ^ (t ^ (t << 13));
    return (s[1] + s[1] + 1) * s[4];
}

// getRandSample (Math.c:304-343) for a list of inLen entries keeping outLen: rank[i] = position of entry i among the
// kept ones, or 0xFFFFFFFF when it is dropped.
static void rand_sample_ranks(uint32_t *state, uint32_t inLen, uint32_t outLen, std::vector<uint8_t> &marked, uint32_t *rank)
{
    marked.assign(inLen, 0);
    bool keepMarked = true;
    uint32_t selectNum = outLen;
    if (outLen > inLen / 2) { keepMarked = false; selectNum = inLen - outLen; }
    for (uint32_t i = inLen - selectNum; i < inLen; i++) {
        const double r = (double)rand_bits(state) / ((double)0xFFFFFFFFu + 1.0);       // getRandDouble, Math.c:289-292
        const uint32_t pos = 0 + (uint32_t)(r * (double)(i + 1 - 0));                   // getRandUInt, Math.c:295-298
        if (marked[pos]) marked[i] = 1; else marked[pos] = 1;
    }
    uint32_t out = 0;
    for (uint32_t i = 0; i < inLen; i++) rank[i] = ((marked[i] != 0) == keepMarked) ? out++ : 0xFFFFFFFFu;
}

extern "C" ya_ctx *ya_open_build(int device, const ya_params *params, const uint8_t *bases, size_t n_base_bytes,
                                 const uint32_t *seq_start, const uint32_t *seq_len, int n_seq,
                                 uint32_t index_max_hits, uint32_t skip_dist)
{
    ya_ctx *c = ya_open_common_for_index(device, params);
    if (!c) return nullptr;
    auto fail = [&](const std::string &m) -> ya_ctx * {
        ya_set_open_error(m); ya_close(c); return nullptr;
    };
    if (!bases || n_seq <= 0 || !seq_start || !seq_len || skip_dist < 1 || index_max_hits < 1) return fail("ya_open_build: bad arguments");
    const int K = params->wordLen;
    if ((int)skip_dist > K) return fail("ya_open_build: skipDist must not exceed wordLen (Main.c:600-604)");
    const size_t n_so = ((size_t)1 << (2 * K)) + 1;
    cudaError_t e;
    if ((e = cudaMalloc(&c->d_bases, n_base_bytes + 64)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_so, n_so * 4)) != cudaSuccess)
        return fail(std::string("cudaMalloc(index): ") + cudaGetErrorString(e));
    cudaMemcpy(c->d_bases, bases, n_base_bytes, cudaMemcpyHostToDevice);
    cudaMemset(c->d_bases + n_base_bytes, 0xEE, 64);
    cudaMemset(c->d_so, 0, n_so * 4);
    c->n_so = n_so; c->n_base_bytes = n_base_bytes;
    c->maxROff = seq_start[n_seq - 1] + seq_len[n_seq - 1];          // BaseSeq.c:121-125
    uint64_t n_pos_total = 0;
    for (int s = 0; s < n_seq; s++) if (seq_len[s] >= (uint32_t)K) n_pos_total += seq_len[s] - K + 1;
    if (n_pos_total >= 0xFFFFFFF0ull) return fail("ya_open_build: reference too large");
    const uint32_t n_keys = (uint32_t)n_pos_total;
    if (c->d_keys0.reserve((size_t)n_keys * 8 + 8) != cudaSuccess || c->d_keys1.reserve((size_t)n_keys * 8 + 8) != cudaSuccess)
        return fail("cudaMalloc(index keys) failed");
    if (c->d_misc.reserve((size_t)n_seq * 4 + 64) != cudaSuccess) return fail("cudaMalloc failed");
    uint32_t *d_firstbad = c->d_misc.as<uint32_t>();
    cudaMemsetAsync(d_firstbad, 0xFF, (size_t)n_seq * 4, c->stream);
    uint64_t *ka = c->d_keys0.as<uint64_t>(), *kb = c->d_keys1.as<uint64_t>();
    uint32_t kbase = 0;
    for (int s = 0; s < n_seq; s++) {
        if (seq_len[s] < (uint32_t)K) continue;
        uint32_t np = seq_len[s] - K + 1;
        if (skip_dist > 1) {
            index_firstbad_kernel<<<(seq_len[s] + 255) / 256, 256, 0, c->stream>>>(c->d_bases, seq_start[s], seq_len[s], d_firstbad + s);
            c->ctr.launches++;
        }
        index_hash_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(c->d_bases, seq_start[s], np, K, skip_dist, d_firstbad + s, c->d_so, ka, kbase);
        kbase += np;
        c->ctr.launches++;
    }
    // counts -> exclusive prefix sums; entry 4^K receives the total (Index.c:185-194)
    uint32_t *d_total = c->d_so + (n_so - 1);
    if (ya_exclusive_scan_u32(c, c->d_so, c->d_so, n_so - 1, d_total) != YA_OK) return fail("index scan failed: " + c->err);
    uint32_t total = 0;
    cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, c->stream);
    // k-mers over the hit cap (Index.c:283-288)
    DevBuf d_over, d_nover;
    uint32_t overCap = 1u << 20, nOver = 0;
    std::vector<OverFull> over;
    for (;;) {
        if (d_over.reserve((size_t)overCap * sizeof(OverFull)) != cudaSuccess || d_nover.reserve(64) != cudaSuccess) return fail("cudaMalloc failed");
        cudaMemsetAsync(d_nover.p, 0, 4, c->stream);
        index_overfull_kernel<<<148 * 8, 256, 0, c->stream>>>(c->d_so, n_so - 1, index_max_hits, d_over.as<OverFull>(), overCap, d_nover.as<uint32_t>());
        c->ctr.launches++;
        cudaMemcpyAsync(&nOver, d_nover.p, 4, cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { d_over.release(); d_nover.release(); return fail("index build: kernel failure"); }
        if (nOver <= overCap) break;
        overCap = nOver + 1024;                                       // (a list longer than the first guess: once more)
    }
    over.resize(nOver);
    if (nOver) cudaMemcpy(over.data(), d_over.p, (size_t)nOver * sizeof(OverFull), cudaMemcpyDeviceToHost);
    d_over.release(); d_nover.release();
    if (ya_radix_sort_u64(c, ka, kb, n_keys, 32, 32 + 2 * K) != YA_OK) return fail("index sort failed: " + c->err);
    {
        // the list of the last k-mer (all G) shares its sort digits with the skipped positions' keys: bring its entries in front
        uint32_t tailStart = 0;
        cudaMemcpyAsync(&tailStart, c->d_so + (n_so - 2), 4, cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("index build: kernel failure");
        const uint32_t nTail = n_keys - tailStart, nValidTail = total - tailStart;
        if (nValidTail > 0 && nTail > nValidTail) {
            DevBuf d_flag, d_idx;
            if (d_flag.reserve((size_t)nTail * 4 + 16) != cudaSuccess || d_idx.reserve((size_t)nTail * 4 + 16) != cudaSuccess) return fail("cudaMalloc failed");
            index_tail_flag_kernel<<<(nTail + 255) / 256, 256, 0, c->stream>>>(ka + tailStart, nTail, d_flag.as<uint32_t>());
            if (ya_exclusive_scan_u32(c, d_flag.as<uint32_t>(), d_idx.as<uint32_t>(), nTail, nullptr) != YA_OK) return fail("index scan failed: " + c->err);
            index_tail_scatter_kernel<<<(nTail + 255) / 256, 256, 0, c->stream>>>(ka + tailStart, nTail, d_flag.as<uint32_t>(), d_idx.as<uint32_t>(), kb + tailStart);
            cudaMemcpyAsync(ka + tailStart, kb + tailStart, (size_t)nValidTail * 8, cudaMemcpyDeviceToDevice, c->stream);
            c->ctr.launches += 2;
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("index build: kernel failure");
            d_flag.release(); d_idx.release();
        }
    }
    if (nOver == 0) {
        if ((e = cudaMalloc(&c->d_roa, ((size_t)total + 8) * 4)) != cudaSuccess)
            return fail(std::string("cudaMalloc(roa): ") + cudaGetErrorString(e));
        if (total) index_take_pos_kernel<<<(total + 255) / 256, 256, 0, c->stream>>>(ka, total, c->d_roa);
        c->ctr.launches++;
        c->n_roa = total;
    } else {
        // pass 3: one xorshift stream through the over-full k-mers in hash order (Index.c:271-315)
        std::sort(over.begin(), over.end(), [](const OverFull &a, const OverFull &b) { return a.hash < b.hash; });
        std::vector<OverPlan> plan(nOver);
        std::vector<uint32_t> soOld(nOver);
        uint64_t nRank = 0;
        for (uint32_t k = 0; k < nOver; k++) nRank += over[k].count;
        if (nRank >= 0xFFFFFFF0ull) return fail("ya_open_build: over-full lists too long");
        std::vector<uint32_t> rank((size_t)nRank);
        std::vector<uint8_t> marked;
        uint32_t state[5] = {123456789u, 362436069u, 521288629u, 88675123u, 886756453u};       // initRandStateDefault, Math.c:252-256
        uint32_t rankOff = 0, dropped = 0;
        for (uint32_t k = 0; k < nOver; k++) {
            rand_sample_ranks(state, over[k].count, index_max_hits, marked, rank.data() + rankOff);
            plan[k].hash = over[k].hash; plan[k].rank_off = rankOff; plan[k].dropped_before = dropped;
            plan[k].dropped = over[k].count - index_max_hits;
            rankOff += over[k].count; dropped += plan[k].dropped;
        }
        // the old starting offsets of the over-full lists (a key's index in its list = its position - that offset)
        DevBuf d_plan, d_rank, d_soOld;
        if (d_plan.reserve((size_t)nOver * sizeof(OverPlan)) != cudaSuccess || d_rank.reserve((size_t)nRank * 4 + 16) != cudaSuccess ||
            d_soOld.reserve((size_t)nOver * 4 + 16) != cudaSuccess) return fail("cudaMalloc(sampling plan) failed");
        for (uint32_t k = 0; k < nOver; k++) soOld[k] = over[k].start;
        cudaMemcpyAsync(d_plan.p, plan.data(), (size_t)nOver * sizeof(OverPlan), cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(d_rank.p, rank.data(), (size_t)nRank * 4, cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(d_soOld.p, soOld.data(), (size_t)nOver * 4, cudaMemcpyHostToDevice, c->stream);
        const uint32_t newTotal = total - dropped;
        if ((e = cudaMalloc(&c->d_roa, ((size_t)newTotal + 8) * 4)) != cudaSuccess)
            return fail(std::string("cudaMalloc(roa): ") + cudaGetErrorString(e));
        index_resample_so_kernel<<<148 * 16, 256, 0, c->stream>>>(c->d_so, n_so, d_plan.as<OverPlan>(), (int)nOver);
        index_resample_roa_kernel<<<(total + 255) / 256, 256, 0, c->stream>>>(ka, total, c->d_so, d_plan.as<OverPlan>(), (int)nOver,
                                                                              d_soOld.as<uint32_t>(), d_rank.as<uint32_t>(), c->d_roa);
        c->ctr.launches += 2;
        cudaStreamSynchronize(c->stream);
        d_plan.release(); d_rank.release(); d_soOld.release();
        c->n_roa = newTotal;
    }
    cudaMemsetAsync(c->d_roa + c->n_roa, 0, 32, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return fail("index build: kernel failure");
    if (ya_build_lowmask(c, bases, n_base_bytes) != YA_OK) return fail("cudaMalloc(lowmask) failed");
    c->d_keys0.release(); c->d_keys1.release();
    return c;
}

extern "C" int ya_index_sizes(const ya_ctx *c, size_t *n_so, size_t *n_roa)
{
    if (!c) return YA_E_ARG;
    if (n_so) *n_so = c->n_so;
    if (n_roa) *n_roa = c->n_roa;
    return YA_OK;
}

extern "C" int ya_index_download(ya_ctx *c, uint32_t *so, uint32_t *roa)
{
    if (!c || !so) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    YA_CUDA(c, cudaMemcpy(so, c->d_so, c->n_so * 4, cudaMemcpyDeviceToHost));
    if (c->n_roa && roa) YA_CUDA(c, cudaMemcpy(roa, c->d_roa, c->n_roa * 4, cudaMemcpyDeviceToHost));
    return YA_OK;
}
