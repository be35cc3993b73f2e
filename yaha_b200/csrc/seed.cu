// seed.cu -- stages 1 and 2 of the alignment hot path on the device.
//
//   K1  seed_count_kernel   k-mer hash + starting-offset gather        (ref: Query.c:233-244, 365-412)
//   K2a expand_hits_kernel  ROA gather -> 64-bit (segment, diagonal, qo) keys
//                                                                     (ref: QueryMatch.c:56-69, 95-96)
//   K2b radix_*             stable LSD radix sort on (segment, diagonal); replaces the binary-heap
//                           k-way merge (QueryMatch.c:70-116, QueryHeap.inl) -- keys are distinct, so
//                           the merge order IS ascending key order
//   K2c frag_*/region_*     head-flag scans: coalesce abutting seeds into fragments
//                           (QueryMatch.c:99-115), cut regions where neighbouring diagonals differ by
//                           more than maxGap (QueryMatch.c:146-158), drop singleton regions shorter than
//                           minMatch (QueryMatch.c:281-290), compact the survivors
//
// A "segment" is one (read, strand) pair: index 2*read + strand.  All of this is HBM-bound
// integer work; no tensor cores.
#include "common.cuh"
#include <algorithm>

#define SEG_SHIFT 47           // key = seg << 47 | diag << 15 | qo
#define QO_BITS   15
#define QO_MASK   0x7FFFu

// ------------------------------------------------------------------------------------------
// K1: one warp per segment.  Each lane takes 4 consecutive query offsets per iteration: 18 code
// bytes give 4 rolling hashes (Query.c:233-244,409-411), and all 8 starting-offset gathers of the
// lane are in flight before the first is consumed (the table is 4 GiB: every probe is a DRAM
// sector miss, so memory-level parallelism is what sets the rate).
// ------------------------------------------------------------------------------------------
#define K1_Q 4
// Random 4-byte gather with the smallest L2 prefetch size PTX offers (64 B): the default on this part
// fills a whole 128 B line per miss, which quadruples the DRAM traffic of a one-word probe.
__device__ __forceinline__ uint32_t gather_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(128)
seed_count_kernel(const uint8_t *__restrict__ fwd, const uint8_t *__restrict__ rev,
                  const uint64_t *__restrict__ read_off, const uint32_t *__restrict__ seg_probe_off,
                  int seg0, int n_seg, int K, uint32_t maxHits,
                  const uint32_t *__restrict__ so, const uint32_t *__restrict__ roa, uint64_t n_roa,
                  const uint32_t *__restrict__ lowmask,
                  uint32_t *__restrict__ cnt, uint32_t *__restrict__ soff,
                  uint32_t *__restrict__ seg_total, uint32_t *__restrict__ seg_eff)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_seg) return;
    const int seg = seg0 + warp;
    const int r = seg >> 1;
    const uint64_t base = read_off[r];
    const int L = (int)(read_off[r + 1] - base);
    const uint8_t *codes = ((seg & 1) ? rev : fwd) + base;
    const uint32_t p0 = seg_probe_off[warp];
    const int m = L - K + 1;                                     // number of probes (Query.c:341)
    const uint32_t mask = 0xFFFFFFFFu >> (32 - 2 * K);
    uint32_t tot = 0, eff = 0;
    for (int q0 = lane * K1_Q; q0 < m; q0 += 32 * K1_Q) {
        // rolling hashes of the K-mers starting at q0 .. q0+3
        uint32_t h = 0;
        uint32_t badmask = 0;                                    // bit j set: code at q0+j is not ACGT
        uint32_t hs[K1_Q];
        bool ok[K1_Q];
        for (int j = 0; j < K - 1; j++) {
            const int pos = q0 + j;
            const uint32_t cde = (pos < L) ? codes[pos] : 4u;
            if (cde > 3) badmask |= 1u << j;
            h = (h << 2) | (cde & 3u);
        }
#pragma unroll
        for (int t = 0; t < K1_Q; t++) {
            const int pos = q0 + K - 1 + t;
            const uint32_t cde = (pos < L) ? codes[pos] : 4u;
            if (cde > 3) badmask |= 1u << (K - 1 + t);
            h = ((h << 2) | (cde & 3u)) & mask;
            hs[t] = h;
        }
        const uint32_t win = (K >= 32) ? 0xFFFFFFFFu : ((1u << K) - 1u);
#pragma unroll
        for (int t = 0; t < K1_Q; t++) ok[t] = (q0 + t < m) && ((badmask >> t) & win) == 0;
        uint32_t s_lo[K1_Q], s_hi[K1_Q];
#pragma unroll
        for (int t = 0; t < K1_Q; t++) {                         // all gathers issued before any use
            s_lo[t] = ok[t] ? gather_u32(so + hs[t]) : 0u;
            s_hi[t] = ok[t] ? gather_u32(so + hs[t] + 1) : 0u;
        }
        uint32_t c_eff[K1_Q], lastHit[K1_Q];
#pragma unroll
        for (int t = 0; t < K1_Q; t++) {
            const uint32_t c = s_hi[t] - s_lo[t];                // Query.c:391
            c_eff[t] = (ok[t] && c <= maxHits) ? c : 0u;         // Query.c:392
            // only k-mers that occur in the first 32 K reference bases can have all hits below qo
            const bool maybeLow = c_eff[t] && ((__ldg(lowmask + ((hs[t] & 0xFFFFFu) >> 5)) >> (hs[t] & 31u)) & 1u);
            lastHit[t] = maybeLow ? gather_u32(roa + s_lo[t] + c_eff[t] - 1) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int t = 0; t < K1_Q; t++) {
            const int qo = q0 + t;
            if (qo >= m) break;
            tot += c_eff[t];
            uint32_t ce = c_eff[t];
            // QueryMatch.c:62-67: if every hit of this k-mer lies below qo the reference keeps
            // reading past the k-mer's list (no newCount < count guard).  Count the extra reads.
            if (ce && lastHit[t] < (uint32_t)qo) {
                uint64_t at = (uint64_t)s_lo[t] + ce;
                while (at < n_roa) {
                    uint32_t x = roa[at++];
                    ce++;
                    if (x >= (uint32_t)qo) break;
                }
            }
            cnt[p0 + qo] = ce;
            soff[p0 + qo] = s_lo[t];
            eff += ce;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, d);
        eff += __shfl_xor_sync(0xffffffffu, eff, d);
    }
    if (lane == 0) { seg_total[warp] = tot; seg_eff[warp] = eff; }
}

// ------------------------------------------------------------------------------------------
// K2a: one warp per segment writes the keys of its probes' hits (ROA order = ascending offset).
// Keys of one segment are generated in ascending qo order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
expand_hits_kernel(const uint32_t *__restrict__ seg_probe_off, int n_seg, int seg_local0,
                   const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ soff,
                   const uint32_t *__restrict__ hit_off, uint32_t probe0,
                   const uint32_t *__restrict__ roa, uint64_t *__restrict__ keys)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_seg) return;
    const uint32_t pa = seg_probe_off[seg_local0 + warp], pb = seg_probe_off[seg_local0 + warp + 1];
    const uint64_t hi_bits = (uint64_t)warp << SEG_SHIFT;
    for (uint32_t gp = pa + lane; gp < pb; gp += 32) {
        const uint32_t c = cnt[gp];
        if (c == 0) continue;
        const uint32_t qo = gp - pa;
        const uint32_t *list = roa + soff[gp];
        uint64_t *out = keys + hit_off[gp - probe0];
        for (uint32_t t = 0; t < c; t++) {
            const uint32_t diag = list[t] - qo;                  // wraps for roff < qo (QueryHeap.inl:70-73)
            out[t] = hi_bits | ((uint64_t)diag << QO_BITS) | qo;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2b (common case): segmented sort in shared memory -- one warp (<= 512 keys) or one block
// (<= 8192 keys) per segment, bitonic network on the full 64-bit key (keys are distinct, so no
// stability is needed).  One HBM read and one write per key instead of one per radix pass.
// ------------------------------------------------------------------------------------------
template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
seg_sort_kernel(uint64_t *__restrict__ keys, const uint32_t *__restrict__ seg_key_off, const uint32_t *__restrict__ seg_ids,
                int n_list)
{
    extern __shared__ uint64_t sh[];
    const int item = blockIdx.x;
    if (item >= n_list) return;
    const uint32_t seg = seg_ids[item];
    const uint32_t a = seg_key_off[seg], b = seg_key_off[seg + 1];
    const int n = (int)(b - a);
    if (n <= 1) return;
    int P = 1; while (P < n) P <<= 1;
    for (int i = threadIdx.x; i < P; i += THREADS) sh[i] = (i < n) ? keys[a + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t x = sh[i], y = sh[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { sh[i] = y; sh[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += THREADS) keys[a + i] = sh[i];
}

__global__ void seg_key_off_kernel(const uint32_t *__restrict__ seg_probe_off, int seg_local0, int n_seg, uint32_t probe0,
                                   const uint32_t *__restrict__ hit_off, uint32_t n_probes, uint32_t n_keys,
                                   uint32_t *__restrict__ seg_key_off)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_seg) return;
    if (s == n_seg) { seg_key_off[s] = n_keys; return; }
    const uint32_t p = seg_probe_off[seg_local0 + s] - probe0;
    seg_key_off[s] = (p < n_probes) ? hit_off[p] : n_keys;
}

// ------------------------------------------------------------------------------------------
// K2b: LSD radix sort, 8-bit digits, stable.  Per pass: block histograms -> scan -> scatter.
// ------------------------------------------------------------------------------------------
#define RS_WARPS      8
#define RS_THREADS    (RS_WARPS * 32)
#define RS_PER_WARP   512
#define RS_TILE       (RS_WARPS * RS_PER_WARP)

__global__ void radix_hist_kernel(const uint64_t *__restrict__ keys, uint32_t n, int shift,
                                  uint32_t *__restrict__ hist, uint32_t n_blocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
    for (uint32_t i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        uint32_t g = base + i;
        if (g < n) atomicAdd(&h[(uint32_t)(keys[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void radix_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, uint32_t n,
                                     int shift, const uint32_t *__restrict__ hist_scanned, uint32_t n_blocks)
{
    __shared__ uint32_t wcount[RS_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const uint32_t wbase = blockIdx.x * RS_TILE + w * RS_PER_WARP;
    // pass 1: per-warp digit counts
    for (int it = 0; it < RS_PER_WARP / 32; it++) {
        uint32_t g = wbase + it * 32 + lane;
        if (g < n) atomicAdd(&wcount[w][(uint32_t)(in[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    // turn counts into starting positions: global digit base + counts of earlier warps
    {
        uint32_t d = threadIdx.x;           // 256 threads, one digit each
        uint32_t run = hist_scanned[d * n_blocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) {
            uint32_t c = wcount[ww][d];
            wcount[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // pass 2: stable placement, 32 keys at a time in index order
    const uint32_t lt = (1u << lane) - 1u;
    for (int it = 0; it < RS_PER_WARP / 32; it++) {
        uint32_t g = wbase + it * 32 + lane;
        bool valid = g < n;
        uint64_t key = valid ? in[g] : 0;
        uint32_t d = valid ? ((uint32_t)(key >> shift) & 255u) : (0x100u + lane);
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t rank = __popc(peers & lt);
        int leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if (valid && lane == leader) {
            pos = wcount[w][d];
            wcount[w][d] = pos + __popc(peers);
        }
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (valid) out[pos + rank] = key;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K2c: fragments, regions, compaction.
// ------------------------------------------------------------------------------------------
struct FragRaw { uint32_t diag; uint16_t sqo; uint16_t seg_lo; uint32_t seg; };   // 12 B

__global__ void frag_flag_kernel(const uint64_t *__restrict__ keys, uint32_t n, int K, uint32_t *__restrict__ flag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t head = 1;
    if (i > 0) {
        uint64_t a = keys[i - 1], b = keys[i];
        // same segment and diagonal, and this seed starts no later than one past the previous
        // seed's last base + 1  (QueryMatch.c:99: nextDiag != curDiag || nextQO > curEQO)
        if ((a >> QO_BITS) == (b >> QO_BITS) && (uint32_t)(b & QO_MASK) <= (uint32_t)(a & QO_MASK) + (uint32_t)K) head = 0;
    }
    flag[i] = head;
}

__global__ void frag_write_kernel(const uint64_t *__restrict__ keys, uint32_t n, int K,
                                  const uint32_t *__restrict__ flag, const uint32_t *__restrict__ fidx,
                                  FragRaw *__restrict__ raw, uint16_t *__restrict__ eqo)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = keys[i];
    uint32_t id = fidx[i] + flag[i] - 1;
    if (flag[i]) {
        FragRaw f;
        f.diag = (uint32_t)(k >> QO_BITS);
        f.sqo = (uint16_t)(k & QO_MASK);
        f.seg = (uint32_t)(k >> SEG_SHIFT);
        f.seg_lo = 0;
        raw[id] = f;
    }
    if (i + 1 == n || flag[i + 1]) eqo[id] = (uint16_t)((k & QO_MASK) + K - 1);
}

// The fragment / region / survivor counts stay on the device: these kernels are launched over the upper bound
// (the number of hits), read the real count from *d_nf and zero their flag beyond it so that the scans can run over
// the bound as well -- no host round trip between the steps.
__global__ void region_flag_kernel(const FragRaw *__restrict__ raw, const uint32_t *__restrict__ d_nf, uint32_t n_upper, uint32_t maxGap,
                                   uint32_t *__restrict__ rflag, uint32_t *__restrict__ seg_first)
{
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nf = *d_nf;
    if (f >= nf) { if (f < n_upper) rflag[f] = 0; return; }
    uint32_t head = 1;
    bool newseg = true;
    if (f > 0) {
        FragRaw a = raw[f - 1], b = raw[f];
        if (a.seg == b.seg) {
            newseg = false;
            uint32_t diff = a.diag > b.diag ? a.diag - b.diag : b.diag - a.diag;    // FragsClumps.inl:133-137
            if (diff <= maxGap) head = 0;
        }
    }
    rflag[f] = head;
    if (newseg) seg_first[raw[f].seg] = f;
}

__global__ void region_start_kernel(const uint32_t *__restrict__ rflag, const uint32_t *__restrict__ ridx,
                                    const uint32_t *__restrict__ d_nf, uint32_t *__restrict__ rstart,
                                    const uint32_t *__restrict__ d_nreg)
{
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nf = *d_nf;
    if (f == 0) rstart[*d_nreg] = nf;
    if (f >= nf) return;
    if (rflag[f]) rstart[ridx[f]] = f;
}

__global__ void keep_flag_kernel(const FragRaw *__restrict__ raw, const uint16_t *__restrict__ eqo,
                                 const uint32_t *__restrict__ rflag, const uint32_t *__restrict__ ridx,
                                 const uint32_t *__restrict__ rstart, const uint32_t *__restrict__ d_nf, uint32_t n_upper,
                                 uint32_t minMatch, uint32_t *__restrict__ keep)
{
    uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= *d_nf) { if (f < n_upper) keep[f] = 0; return; }
    uint32_t rid = ridx[f] + rflag[f] - 1;
    uint32_t members = rstart[rid + 1] - rstart[rid];
    uint32_t refLen = (uint32_t)eqo[f] - raw[f].sqo + 1;
    keep[f] = (members > 1 || refLen >= minMatch) ? 1u : 0u;        // QueryMatch.c:281-290
}

__global__ void compact_kernel(const FragRaw *__restrict__ raw, const uint16_t *__restrict__ eqo,
                               const uint32_t *__restrict__ rflag, const uint32_t *__restrict__ ridx,
                               const uint32_t *__restrict__ keep, const uint32_t *__restrict__ kidx,
                               const uint32_t *__restrict__ seg_first, const uint32_t *__restrict__ d_nf,
                               ya_frag *__restrict__ out, uint32_t *__restrict__ region_out,
                               ya_strand_frags *__restrict__ strands)
{
    // The strand counters are updated once per warp and strand (lanes of one strand elect a leader): a repeat-rich strand
    // holds 10^5 fragments, and one atomic per fragment on its record serialised the whole kernel (cfg5: 4.7 ms per launch, a third of the batch's kernel time).
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = f < *d_nf;
    const int lane = threadIdx.x & 31;
    FragRaw r; r.diag = 0; r.sqo = 0; r.seg_lo = 0; r.seg = 0xFFFFFFFFu;
    if (valid) r = raw[f];
    const bool kept = valid && keep[f];
    uint32_t o = 0xFFFFFFFFu;
    if (kept) {
        o = kidx[f];
        ya_frag g;
        g.startRefOff = r.diag + r.sqo;
        g.startQueryOff = r.sqo;
        g.endQueryOff = eqo[f];
        g.hitCount = 0;
        g.refLen = (uint16_t)(g.endQueryOff - g.startQueryOff + 1);      // FragsClumps.inl:44-46
        out[o] = g;
        const uint32_t f0 = seg_first[r.seg];
        region_out[o] = (ridx[f] + rflag[f] - 1) - (ridx[f0] + rflag[f0] - 1);
    }
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, r.seg);
    const uint32_t keptMask = __ballot_sync(0xFFFFFFFFu, kept) & peers;
    if (valid && lane == __ffs(peers) - 1) {
        atomicAdd(&strands[r.seg].n_frags_all, (uint32_t)__popc(peers));
        if (keptMask) atomicAdd(&strands[r.seg].n_frags, (uint32_t)__popc(keptMask));
    }
    // (survivors keep the fragments' order, so the lowest kept lane of a strand holds the warp's smallest output index)
    if (kept && lane == __ffs(keptMask) - 1) atomicMin(&strands[r.seg].first, o);
}

__global__ void init_strands_kernel(ya_strand_frags *s, const uint32_t *__restrict__ seg_total, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ya_strand_frags v; v.first = 0xFFFFFFFFu; v.n_frags = 0; v.n_frags_all = 0; v.total_hits = seg_total[i];
    s[i] = v;
}

// ------------------------------------------------------------------------------------------
// K2, fused (the common case: every segment of the chunk has at most 8192 hits).  One warp (<= 512 hits) or one block per
// segment does the whole of stage 2 in shared memory: expand the probes' ROA lists into (diagonal, qo) keys, sort them
// (bitonic; keys are distinct), coalesce abutting seeds into fragments (QueryMatch.c:99-115), cut regions where neighbouring
// diagonals differ by more than maxGap (QueryMatch.c:146-158), drop singleton regions shorter than minMatch
// (QueryMatch.c:281-290) and write the survivors to the segment's place in a staging array.  HBM traffic: 4 B per hit read,
// 16 B per survivor written -- against ~100 B per hit of the expand / sort / flag / scan / write chain it replaces, and two
// launches (+ one compaction) instead of twenty.
// ------------------------------------------------------------------------------------------
template <int THREADS>
__device__ __forceinline__ uint32_t k2_block_excl_scan(uint32_t v, uint32_t *total, uint32_t *wsum)
{
    // exclusive scan of one value per thread over the block (THREADS = 32: a single warp)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (THREADS == 32) { *total = __shfl_sync(0xffffffffu, inc, 31); return inc - v; }
    __syncthreads();                                   // (wsum may still be read from the previous scan)
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t x = (lane < THREADS / 32) ? wsum[lane] : 0u;
        uint32_t xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += t; }
        if (lane < THREADS / 32) wsum[lane] = xi - x;
        if (lane == 31) wsum[32] = xi;
    }
    __syncthreads();
    *total = wsum[32];
    return inc - v + wsum[w];
}

struct K2Frag { uint32_t diag; uint16_t sqo, eqo; };                 // 8 B: overlays the sorted keys

template <int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
k2_fused_kernel(const uint32_t *__restrict__ seg_ids, int n_list, const uint32_t *__restrict__ seg_probe_off, int seg_local0,
                const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ soff, const uint32_t *__restrict__ roa,
                const uint32_t *__restrict__ seg_key_off, int K, uint32_t maxGap, uint32_t minMatch,
                ya_frag *__restrict__ stage_frags, uint32_t *__restrict__ stage_region, uint32_t *__restrict__ seg_nall,
                uint32_t *__restrict__ seg_nkeep, uint32_t *__restrict__ totals)
{
    constexpr int IPT = CAP / THREADS;                               // items per thread, blocked: thread t owns [t*IPT, (t+1)*IPT)
    static_assert(IPT <= 32, "per-thread flags are kept in 32-bit masks");
    extern __shared__ uint64_t sh[];                                 // CAP keys, later CAP fragment records
    __shared__ uint32_t wsum[33];
    const int item = blockIdx.x;
    if (item >= n_list) return;
    const uint32_t seg = seg_ids[item];
    const uint32_t base = seg_key_off[seg];
    const int n = (int)(seg_key_off[seg + 1] - base);
    const uint32_t pa = seg_probe_off[seg_local0 + seg], pb = seg_probe_off[seg_local0 + seg + 1];
    const int t = threadIdx.x;
    // ---- expand: probes in qo order, each probe's list in ROA order (ascending offset)
    {
        uint32_t run = 0;                                            // hits of the probes handled in earlier rounds
        for (uint32_t p0 = pa; p0 < pb; p0 += THREADS) {
            const uint32_t gp = p0 + t;
            const uint32_t c = (gp < pb) ? cnt[gp] : 0u;
            uint32_t tot;
            const uint32_t at = run + k2_block_excl_scan<THREADS>(c, &tot, wsum);
            if (c) {
                const uint32_t qo = gp - pa;
                const uint32_t *list = roa + soff[gp];
                for (uint32_t k = 0; k < c; k++) {
                    const uint32_t diag = list[k] - qo;              // wraps for roff < qo (QueryHeap.inl:70-73)
                    sh[at + k] = ((uint64_t)diag << QO_BITS) | qo;
                }
            }
            run += tot;
        }
    }
    int P = 1; while (P < n) P <<= 1;
    for (int i = n + t; i < P; i += THREADS) sh[i] = ~0ull;
    __syncthreads();
    // ---- sort
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < P; i += THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t x = sh[i], y = sh[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { sh[i] = y; sh[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    // ---- fragments: a key opens one unless it continues the previous key's (QueryMatch.c:99: nextDiag != curDiag || nextQO > curEQO)
    uint64_t key[IPT];
    uint32_t headMask = 0, tailMask = 0;                             // bit k: item t*IPT + k opens / closes a fragment
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int i = t * IPT + k;
        key[k] = 0;
        if (i < n) {
            const uint64_t b = sh[i];
            key[k] = b;
            bool head = true, tail = true;
            if (i > 0) {
                const uint64_t a = sh[i - 1];
                if ((a >> QO_BITS) == (b >> QO_BITS) && (uint32_t)(b & QO_MASK) <= (uint32_t)(a & QO_MASK) + (uint32_t)K) head = false;
            }
            if (i + 1 < n) {
                const uint64_t c2 = sh[i + 1];
                if ((c2 >> QO_BITS) == (b >> QO_BITS) && (uint32_t)(c2 & QO_MASK) <= (uint32_t)(b & QO_MASK) + (uint32_t)K) tail = false;
            }
            headMask |= (head ? 1u : 0u) << k; tailMask |= (tail ? 1u : 0u) << k;
        }
    }
    uint32_t nf;
    uint32_t fid = k2_block_excl_scan<THREADS>((uint32_t)__popc(headMask), &nf, wsum);   // id of this thread's first head
    __syncthreads();                                                 // every key is in registers: the array becomes fragment records
    K2Frag *fr = reinterpret_cast<K2Frag *>(sh);
    // (a fragment's last key may belong to another thread than its first: `fid - 1` there is that thread's running id, which
    //  counts the heads at or before the key in key order -- the fragment this key closes)
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        if ((headMask >> k) & 1u) { fr[fid].diag = (uint32_t)(key[k] >> QO_BITS); fr[fid].sqo = (uint16_t)(key[k] & QO_MASK); fid++; }
        if ((tailMask >> k) & 1u) fr[fid - 1].eqo = (uint16_t)((key[k] & QO_MASK) + K - 1);
    }
    __syncthreads();
    // ---- regions and survivors: fragments blocked over the threads the same way.  A region is cut where neighbouring
    // diagonals differ by more than maxGap (QueryMatch.c:146-158); a fragment whose region has no second member survives only
    // with refLen >= minMatch (QueryMatch.c:281-290).
    uint32_t rheadMask = 0, keepMask = 0;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const uint32_t f = (uint32_t)(t * IPT + k);
        if (f < nf) {
            const uint32_t d = fr[f].diag;
            bool rhead = true, rlast = true;
            if (f > 0) { const uint32_t a = fr[f - 1].diag; if ((a > d ? a - d : d - a) <= maxGap) rhead = false; }       // FragsClumps.inl:133-137
            if (f + 1 < nf) { const uint32_t b = fr[f + 1].diag; if ((b > d ? b - d : d - b) <= maxGap) rlast = false; }
            const uint32_t refLen = (uint32_t)fr[f].eqo - fr[f].sqo + 1;
            rheadMask |= (rhead ? 1u : 0u) << k;
            keepMask |= ((!(rhead && rlast) || refLen >= minMatch) ? 1u : 0u) << k;
        }
    }
    uint32_t nreg, nk;
    const uint32_t rid0 = k2_block_excl_scan<THREADS>((uint32_t)__popc(rheadMask), &nreg, wsum);
    uint32_t kid = k2_block_excl_scan<THREADS>((uint32_t)__popc(keepMask), &nk, wsum);
    (void)nreg;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        if ((keepMask >> k) & 1u) {
            const uint32_t f = (uint32_t)(t * IPT + k);
            ya_frag g;
            g.startRefOff = fr[f].diag + fr[f].sqo;
            g.startQueryOff = fr[f].sqo;
            g.endQueryOff = fr[f].eqo;
            g.hitCount = 0;
            g.refLen = (uint16_t)(g.endQueryOff - g.startQueryOff + 1);   // FragsClumps.inl:44-46
            stage_frags[base + kid] = g;
            stage_region[base + kid] = rid0 + (uint32_t)__popc(rheadMask & (0xFFFFFFFFu >> (31 - k))) - 1u;   // region ordinal within the strand
            kid++;
        }
    }
    if (t == 0) { seg_nall[seg] = nf; seg_nkeep[seg] = nk; atomicAdd(&totals[0], nf); atomicAdd(&totals[2], nk); }
}

// survivors of every segment from the staging array (at the segment's key offset) to their final, dense place; strand records
__global__ void __launch_bounds__(128)
k2_compact_kernel(int n_seg, const uint32_t *__restrict__ seg_key_off, const uint32_t *__restrict__ seg_nall,
                  const uint32_t *__restrict__ seg_nkeep, const uint32_t *__restrict__ keep_off, const uint32_t *__restrict__ seg_total,
                  const ya_frag *__restrict__ stage_frags, const uint32_t *__restrict__ stage_region,
                  ya_frag *__restrict__ out, uint32_t *__restrict__ region_out, ya_strand_frags *__restrict__ strands)
{
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= n_seg) return;
    const uint32_t nk = seg_nkeep[warp], from = seg_key_off[warp], to = keep_off[warp];
    for (uint32_t k = lane; k < nk; k += 32) { out[to + k] = stage_frags[from + k]; region_out[to + k] = stage_region[from + k]; }
    if (lane == 0) {
        ya_strand_frags v;
        v.first = nk ? to : 0xFFFFFFFFu; v.n_frags = nk; v.n_frags_all = seg_nall[warp]; v.total_hits = seg_total[warp];
        strands[warp] = v;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int ya_radix_sort_u64(ya_ctx *c, uint64_t *&a, uint64_t *&b, uint32_t n, int lo_bit, int hi_bit)
{
    uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
    YA_CUDA(c, c->d_hist.reserve((size_t)256 * n_blocks * 4));
    uint32_t *hist = c->d_hist.as<uint32_t>();
    for (int shift = lo_bit; shift < hi_bit; shift += 8) {
        radix_hist_kernel<<<n_blocks, RS_THREADS, 0, c->stream>>>(a, n, shift, hist, n_blocks);
        int rc = ya_exclusive_scan_u32(c, hist, hist, (size_t)256 * n_blocks, nullptr);
        if (rc != YA_OK) return rc;
        radix_scatter_kernel<<<n_blocks, RS_THREADS, 0, c->stream>>>(a, b, n, shift, hist, n_blocks);
        c->ctr.launches += 2;
        std::swap(a, b);
    }
    YA_CUDA(c, cudaGetLastError());
    return YA_OK;
}

// device_only (ya_align_batch): the survivors stay on the device, nothing but the counts comes back; `out` only receives
// n_frags.  Needs the whole batch in one chunk (YA_E_STATE otherwise: the caller takes the call-by-call path).
int ya_seed_frags_impl(ya_ctx *c, ya_frag_batch *out, bool device_only)
{
    if (!c || !out || (!device_only && (!out->strands || (out->frags_cap && (!out->frags || !out->region))))) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    const int n_reads = c->n_reads;
    const int n_seg = 2 * n_reads;
    const int K = c->P.wordLen;
    out->n_frags = 0; out->frags_needed = 0;
    if (n_reads == 0) return YA_OK;
    cudaStream_t st = c->stream;
    AllocScope allocScope(st);

    // probe offsets per segment
    std::vector<uint32_t> h_po((size_t)n_seg + 1);
    uint64_t acc = 0;
    for (int s = 0; s < n_seg; s++) {
        h_po[s] = (uint32_t)acc;
        int64_t L = (int64_t)(c->h_read_off[(s >> 1) + 1] - c->h_read_off[s >> 1]);
        if (L >= K) acc += (uint64_t)(L - K + 1);
        if (acc >= 0xFFFF0000ull) return ya_fail(c, YA_E_ARG, "batch too large: more than 2^32 k-mer probes");
    }
    h_po[n_seg] = (uint32_t)acc;
    const uint32_t n_probes = (uint32_t)acc;
    DeviceTurn turn(c->device);                 // held to the end of the call (results are copied back last)
    YA_CUDA(c, c->d_seg_probe_off.reserve(((size_t)n_seg + 1) * 4));
    YA_CUDA(c, c->d_cnt.reserve((size_t)n_probes * 4 + 16));
    YA_CUDA(c, c->d_soff.reserve((size_t)n_probes * 4 + 16));
    YA_CUDA(c, c->d_misc.reserve((size_t)n_seg * 8 + 64));
    YA_CUDA(c, c->d_strand_out.reserve((size_t)n_seg * sizeof(ya_strand_frags)));
    YA_CUDA(c, c->h_stage.reserve((size_t)n_seg * 8 + 64));
    uint32_t *d_po = c->d_seg_probe_off.as<uint32_t>();
    uint32_t *d_cnt = c->d_cnt.as<uint32_t>(), *d_soff = c->d_soff.as<uint32_t>();
    uint32_t *d_seg_total = c->d_misc.as<uint32_t>(), *d_seg_eff = d_seg_total + n_seg;
    ya_strand_frags *d_strands = c->d_strand_out.as<ya_strand_frags>();

    YA_CUDA(c, cudaEventRecord(c->ev[0], st));
    YA_CUDA(c, cudaMemcpyAsync(d_po, h_po.data(), ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, st));
    {
        int threads = 128, warps_per_block = threads / 32;
        int blocks = (n_seg + warps_per_block - 1) / warps_per_block;
        YA_CUDA(c, cudaEventRecord(c->ev[3], st));
        seed_count_kernel<<<blocks, threads, 0, st>>>(c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(),
                                                     c->d_read_off.as<uint64_t>(), d_po, 0, n_seg, K,
                                                     (uint32_t)c->P.maxHits, c->d_so, c->d_roa, (uint64_t)c->n_roa, c->d_lowmask,
                                                     d_cnt, d_soff, d_seg_total, d_seg_eff);
        YA_CUDA(c, cudaEventRecord(c->ev[4], st));
        init_strands_kernel<<<(n_seg + 255) / 256, 256, 0, st>>>(d_strands, d_seg_total, n_seg);
        c->ctr.launches += 2;
    }
    uint32_t *h_seg_eff = c->h_stage.as<uint32_t>();
    YA_CUDA(c, cudaMemcpyAsync(h_seg_eff, d_seg_eff, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, ya_stream_wait(st));
    c->ctr.probes += n_probes;
    { float msl = 0; cudaEventElapsedTime(&msl, c->ev[3], c->ev[4]); c->ctr.ms_lookup += msl; }

    // chunk plan: <= 2^16 reads and <= MAX_KEYS hits per chunk (segment id fits 17 bits)
    const uint64_t MAX_KEYS = 1ull << 28;
    size_t out_base = 0;          // survivors written so far (device-side running offset)
    c->seed_chunks = 0; c->seed_nkeep = 0;
    bool overflow = false;
    int s0 = 0;
    while (s0 < n_seg) {
        int s1 = s0;
        uint64_t hits = 0;
        while (s1 < n_seg && (s1 - s0) < (1 << 17)) {
            uint64_t add = (uint64_t)h_seg_eff[s1] + h_seg_eff[s1 + 1];
            if (s1 > s0 && hits + add > MAX_KEYS) break;
            hits += add; s1 += 2;
        }
        if (hits >= 0xFFFFFFF0ull) return ya_fail(c, YA_E_ARG, "a single read produces more than 2^32 seed hits");
        if (device_only && (s0 != 0 || s1 != n_seg)) return ya_fail(c, YA_E_STATE, "batch needs several seed chunks");
        const uint32_t n_keys = (uint32_t)hits;
        const int cseg = s1 - s0;
        c->ctr.hits += n_keys;
        if (n_keys > 0) {
            uint32_t maxSeg = 0;
            for (int sgi = s0; sgi < s1; sgi++) maxSeg = std::max(maxSeg, h_seg_eff[sgi]);
            const bool noFused = getenv("YA_SEED_CHAIN") != nullptr || getenv("YA_SEED_RADIX") != nullptr;   // (tests: the kernel chain)
            const size_t NU = n_keys;
            uint32_t *d_tot = c->d_misc.as<uint32_t>() + 2 * (size_t)n_seg;    // 4 spare words: fragments, regions, survivors
            if (maxSeg <= 8192 && !noFused) {
                // ---- fused stage 2: one warp / block per segment in shared memory (k2_fused_kernel), then one compaction
                YA_CUDA(c, c->d_frags_out.reserve(NU * sizeof(ya_frag) + 16));
                YA_CUDA(c, c->d_region_out.reserve(NU * 4 + 16));
                YA_CUDA(c, c->d_frags_all.reserve(NU * sizeof(ya_frag) + 16));          // staging: survivors at their segment's key offset
                YA_CUDA(c, c->d_regidx.reserve(NU * 4 + 16));
                YA_CUDA(c, c->d_regstart.reserve(((size_t)6 * cseg + 8) * 4 + 64));
                std::vector<uint32_t> &small = c->seed_small, &big = c->seed_big, &koff = c->seed_koff;
                small.clear(); big.clear(); koff.resize((size_t)cseg + 1);
                uint32_t run = 0;
                for (int sgi = s0; sgi < s1; sgi++) {
                    const uint32_t e = h_seg_eff[sgi];
                    koff[(size_t)(sgi - s0)] = run; run += e;
                    if (e >= 1 && e <= 512) small.push_back((uint32_t)(sgi - s0));
                    else if (e > 512) big.push_back((uint32_t)(sgi - s0));
                }
                koff[(size_t)cseg] = run;
                uint32_t *d_sko = c->d_regstart.as<uint32_t>();
                uint32_t *d_nall = d_sko + cseg + 1, *d_nkeep = d_nall + cseg, *d_koff = d_nkeep + cseg;
                uint32_t *d_small = d_koff + cseg + 1, *d_big = d_small + small.size();
                YA_CUDA(c, cudaMemcpyAsync(d_sko, koff.data(), ((size_t)cseg + 1) * 4, cudaMemcpyHostToDevice, st));
                YA_CUDA(c, cudaMemsetAsync(d_nall, 0, (size_t)2 * cseg * 4, st));
                YA_CUDA(c, cudaMemsetAsync(d_tot, 0, 16, st));
                ya_frag *stage = c->d_frags_all.as<ya_frag>();
                uint32_t *stage_reg = c->d_regidx.as<uint32_t>();
                if (!small.empty()) {
                    YA_CUDA(c, cudaMemcpyAsync(d_small, small.data(), small.size() * 4, cudaMemcpyHostToDevice, st));
                    k2_fused_kernel<512, 32><<<(unsigned)small.size(), 32, 512 * 8, st>>>(d_small, (int)small.size(), d_po, s0, d_cnt, d_soff, c->d_roa,
                        d_sko, K, (uint32_t)c->P.maxGap, (uint32_t)c->P.minMatch, stage, stage_reg, d_nall, d_nkeep, d_tot);
                    c->ctr.launches++;
                }
                if (!big.empty()) {
                    // (the opt-in to 64 KB of dynamic shared memory is per device: made once per context, which is bound to one)
                    if (!c->big_sort_attr) {
                        YA_CUDA(c, cudaFuncSetAttribute(k2_fused_kernel<8192, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
                        YA_CUDA(c, cudaFuncSetAttribute(seg_sort_kernel<8192, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
                        c->big_sort_attr = true;
                    }
                    YA_CUDA(c, cudaMemcpyAsync(d_big, big.data(), big.size() * 4, cudaMemcpyHostToDevice, st));
                    k2_fused_kernel<8192, 256><<<(unsigned)big.size(), 256, 8192 * 8, st>>>(d_big, (int)big.size(), d_po, s0, d_cnt, d_soff, c->d_roa,
                        d_sko, K, (uint32_t)c->P.maxGap, (uint32_t)c->P.minMatch, stage, stage_reg, d_nall, d_nkeep, d_tot);
                    c->ctr.launches++;
                }
                int rc = ya_exclusive_scan_u32(c, d_nkeep, d_koff, (size_t)cseg, nullptr);
                if (rc != YA_OK) return rc;
                k2_compact_kernel<<<(cseg + 3) / 4, 128, 0, st>>>(cseg, d_sko, d_nall, d_nkeep, d_koff, d_seg_total + s0, stage, stage_reg,
                    c->d_frags_out.as<ya_frag>(), c->d_region_out.as<uint32_t>(), d_strands + s0);
                c->ctr.launches++;
                YA_CUDA(c, cudaGetLastError());
            } else {
            const uint32_t probe0 = h_po[s0], cprobes = h_po[s1] - h_po[s0];
            YA_CUDA(c, c->d_hit_off.reserve((size_t)cprobes * 4 + 16));
            YA_CUDA(c, c->d_keys0.reserve((size_t)n_keys * 8));
            YA_CUDA(c, c->d_keys1.reserve((size_t)n_keys * 8));
            YA_CUDA(c, c->d_regstart.reserve(std::max<size_t>(((size_t)3 * cseg + 1) * 4 + 64, ((size_t)n_keys + 2) * 4)));
            uint32_t *d_hit_off = c->d_hit_off.as<uint32_t>();
            int rc = ya_exclusive_scan_u32(c, d_cnt + probe0, d_hit_off, cprobes, nullptr);
            if (rc != YA_OK) return rc;
            uint64_t *ka = c->d_keys0.as<uint64_t>(), *kb = c->d_keys1.as<uint64_t>();
            expand_hits_kernel<<<(cseg + 3) / 4, 128, 0, st>>>(d_po, cseg, s0, d_cnt, d_soff, d_hit_off, probe0, c->d_roa, ka);
            c->ctr.launches++;
            // sort: segmented shared-memory sort when every segment fits a block, else the global radix sort
            if (maxSeg <= 8192 && !getenv("YA_SEED_RADIX")) {
                std::vector<uint32_t> &small = c->seed_small, &big = c->seed_big;
                small.clear(); big.clear();
                for (int sgi = s0; sgi < s1; sgi++) {
                    const uint32_t e = h_seg_eff[sgi];
                    if (e >= 2 && e <= 512) small.push_back((uint32_t)(sgi - s0));
                    else if (e > 512) big.push_back((uint32_t)(sgi - s0));
                }
                // (d_regstart was sized at the top of the chunk for the id lists here and the region starts later)
                uint32_t *d_sko = c->d_regstart.as<uint32_t>();
                uint32_t *d_small = d_sko + cseg + 1, *d_big = d_small + small.size();
                seg_key_off_kernel<<<(cseg + 256) / 256, 256, 0, st>>>(d_po, s0, cseg, probe0, d_hit_off, cprobes, n_keys, d_sko);
                c->ctr.launches++;
                if (!small.empty()) {
                    YA_CUDA(c, cudaMemcpyAsync(d_small, small.data(), small.size() * 4, cudaMemcpyHostToDevice, st));
                    seg_sort_kernel<512, 32><<<(unsigned)small.size(), 32, 512 * 8, st>>>(ka, d_sko, d_small, (int)small.size());
                    c->ctr.launches++;
                }
                if (!big.empty()) {
                    // (the opt-in to 64 KB of dynamic shared memory is per device: made once per context, which is bound to one)
                    if (!c->big_sort_attr) {
                        YA_CUDA(c, cudaFuncSetAttribute(seg_sort_kernel<8192, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
                        c->big_sort_attr = true;
                    }
                    YA_CUDA(c, cudaMemcpyAsync(d_big, big.data(), big.size() * 4, cudaMemcpyHostToDevice, st));
                    seg_sort_kernel<8192, 256><<<(unsigned)big.size(), 256, 8192 * 8, st>>>(ka, d_sko, d_big, (int)big.size());
                    c->ctr.launches++;
                }
                // (no wait: a copy from pageable memory has left the host vectors when cudaMemcpyAsync returns)
            } else {
                int segbits = 1; while ((1 << segbits) < cseg) segbits++;
                rc = ya_radix_sort_u64(c, ka, kb, n_keys, QO_BITS, SEG_SHIFT + segbits);
                if (rc != YA_OK) return rc;
            }

            // fragments, regions, survivors: every array is sized by the number of hits (>= fragments >= regions,
            // survivors), the counts stay on the device (d_tot[0..2]) and come back with the results
            YA_CUDA(c, c->d_fragflag.reserve(NU * 4));
            YA_CUDA(c, c->d_fragidx.reserve(std::max<size_t>(NU, (size_t)cseg) * 4));
            YA_CUDA(c, c->d_frags_all.reserve(NU * sizeof(FragRaw)));
            YA_CUDA(c, c->d_frag_seg.reserve(NU * 2 + 16));
            YA_CUDA(c, c->d_regflag.reserve(NU * 4));
            YA_CUDA(c, c->d_regidx.reserve(NU * 4));
            YA_CUDA(c, c->d_keep.reserve(NU * 4));
            YA_CUDA(c, c->d_keepidx.reserve(NU * 4));
            YA_CUDA(c, c->d_frags_out.reserve(NU * sizeof(ya_frag) + 16));
            YA_CUDA(c, c->d_region_out.reserve(NU * 4 + 16));
            uint32_t *fflag = c->d_fragflag.as<uint32_t>(), *fidx = c->d_fragidx.as<uint32_t>();
            uint32_t nb = (n_keys + 255) / 256;
            frag_flag_kernel<<<nb, 256, 0, st>>>(ka, n_keys, K, fflag);
            c->ctr.launches++;
            rc = ya_exclusive_scan_u32(c, fflag, fidx, n_keys, d_tot);
            if (rc != YA_OK) return rc;
            FragRaw *raw = c->d_frags_all.as<FragRaw>();
            uint16_t *eqo = c->d_frag_seg.as<uint16_t>();
            frag_write_kernel<<<nb, 256, 0, st>>>(ka, n_keys, K, fflag, fidx, raw, eqo);
            c->ctr.launches++;
            uint32_t *rflag = c->d_regflag.as<uint32_t>(), *ridx = c->d_regidx.as<uint32_t>();
            uint32_t *keep = c->d_keep.as<uint32_t>(), *kidx = c->d_keepidx.as<uint32_t>();
            // seg_first lives in d_hit_off (no longer needed once keys are expanded); the region starts get their own array
            YA_CUDA(c, c->d_hit_off.reserve((size_t)std::max<size_t>(cprobes, cseg) * 4 + 16));
            uint32_t *seg_first = c->d_hit_off.as<uint32_t>();
            uint32_t *rstart = c->d_regstart.as<uint32_t>();       // (the sort's id lists that lived here are consumed by now)
            region_flag_kernel<<<nb, 256, 0, st>>>(raw, d_tot, n_keys, (uint32_t)c->P.maxGap, rflag, seg_first);
            c->ctr.launches++;
            rc = ya_exclusive_scan_u32(c, rflag, ridx, n_keys, d_tot + 1);
            if (rc != YA_OK) return rc;
            region_start_kernel<<<nb, 256, 0, st>>>(rflag, ridx, d_tot, rstart, d_tot + 1);
            keep_flag_kernel<<<nb, 256, 0, st>>>(raw, eqo, rflag, ridx, rstart, d_tot, n_keys, (uint32_t)c->P.minMatch, keep);
            c->ctr.launches += 2;
            rc = ya_exclusive_scan_u32(c, keep, kidx, n_keys, d_tot + 2);
            if (rc != YA_OK) return rc;
            compact_kernel<<<nb, 256, 0, st>>>(raw, eqo, rflag, ridx, keep, kidx, seg_first, d_tot,
                                               c->d_frags_out.as<ya_frag>(), c->d_region_out.as<uint32_t>(),
                                               d_strands + s0);
            c->ctr.launches++;
            YA_CUDA(c, cudaGetLastError());
            }
            // results: the counts, the strand records and -- speculatively -- the first survivors (32 per read)
            uint32_t *h_cnt = c->h_stage.as<uint32_t>() + (size_t)n_seg * 2;          // 3 words after the per-segment totals
            YA_CUDA(c, cudaMemcpyAsync(h_cnt, d_tot, 12, cudaMemcpyDeviceToHost, st));
            if (!device_only)
                YA_CUDA(c, cudaMemcpyAsync(out->strands + s0, d_strands + s0, (size_t)cseg * sizeof(ya_strand_frags),
                                           cudaMemcpyDeviceToHost, st));
            const size_t room = (!device_only && !overflow && out->frags_cap > out_base) ? out->frags_cap - out_base : 0;
            const size_t guess = std::min<size_t>(std::min<size_t>(room, NU), (size_t)16 * (size_t)cseg + 64);
            if (guess) {
                YA_CUDA(c, cudaMemcpyAsync(out->frags + out_base, c->d_frags_out.p, guess * sizeof(ya_frag), cudaMemcpyDeviceToHost, st));
                YA_CUDA(c, cudaMemcpyAsync(out->region + out_base, c->d_region_out.p, guess * 4, cudaMemcpyDeviceToHost, st));
            }
            YA_CUDA(c, ya_stream_wait(st));
            const uint32_t nf = h_cnt[0], nkeep = h_cnt[2];
            c->ctr.frags_all += nf;
            c->ctr.frags_out += nkeep;
            if (device_only) {
            } else if (!overflow && out_base + nkeep <= out->frags_cap) {
                if ((size_t)nkeep > guess) {                              // the speculative copy was too short: fetch the rest
                    YA_CUDA(c, cudaMemcpyAsync(out->frags + out_base + guess, c->d_frags_out.as<ya_frag>() + guess,
                                               ((size_t)nkeep - guess) * sizeof(ya_frag), cudaMemcpyDeviceToHost, st));
                    YA_CUDA(c, cudaMemcpyAsync(out->region + out_base + guess, c->d_region_out.as<uint32_t>() + guess,
                                               ((size_t)nkeep - guess) * 4, cudaMemcpyDeviceToHost, st));
                    YA_CUDA(c, ya_stream_wait(st));
                }
            } else overflow = true;
            // strands of this chunk come back with chunk-relative `first`; fix up on the host
            for (int s = s0; !device_only && s < s1; s++) {
                ya_strand_frags &v = out->strands[s];
                v.first = v.n_frags ? (uint32_t)(v.first + out_base) : (uint32_t)out_base;
            }
            out_base += nkeep;
        } else if (!device_only) {
            YA_CUDA(c, cudaMemcpyAsync(out->strands + s0, d_strands + s0, (size_t)cseg * sizeof(ya_strand_frags),
                                       cudaMemcpyDeviceToHost, st));
            YA_CUDA(c, ya_stream_wait(st));
            for (int s = s0; s < s1; s++) out->strands[s].first = (uint32_t)out_base;
        }
        c->seed_chunks++;
        s0 = s1;
    }
    c->seed_nkeep = out_base;
    YA_CUDA(c, cudaEventRecord(c->ev[1], st));
    YA_CUDA(c, cudaEventSynchronize(c->ev[1]));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->ctr.ms_seed += ms;
    if (overflow) {
        out->frags_needed = out_base;
        return ya_fail(c, YA_E_CAPACITY, "frag output buffer too small");
    }
    out->n_frags = out_base;
    return YA_OK;
}

extern "C" int ya_seed_frags(ya_ctx *c, ya_frag_batch *out) { return ya_seed_frags_impl(c, out, false); }
