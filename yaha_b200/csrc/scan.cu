// scan.cu -- device-wide exclusive prefix sum over uint32 (reduce / recurse / down-sweep).
// Used for hit offsets, head-flag -> fragment ids, region ids and output compaction
// (the "scan" half of north_star stage 2).  HBM-bound: 2 reads + 1 write of 4 B per item.
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS   8
#define SCAN_TILE    (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0;
        uint32_t si = warp_incl_scan(s);
        if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) *total = si;
    }
    __syncthreads();
    uint32_t r = inc - v + wsum[w];
    return r;
}

__global__ void scan_tile_sums(const uint32_t *__restrict__ in, uint32_t *__restrict__ sums, size_t n)
{
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    __shared__ uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void scan_downsweep(const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                               const uint32_t *__restrict__ tile_off, size_t n)
{
    // each thread owns SCAN_ITEMS consecutive items so that the scan order is the index order
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + k;
        v[k] = (i < n) ? in[i] : 0;
        s += v[k];
    }
    __shared__ uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + (tile_off ? tile_off[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        size_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

__global__ void scan_write_total(const uint32_t *__restrict__ last_in, const uint32_t *__restrict__ out,
                                 size_t n, uint32_t *total)
{
    *total = out[n - 1] + *last_in;
}

static int scan_rec(ya_ctx *c, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *tmp, size_t tmp_words)
{
    if (n == 0) return YA_OK;
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_downsweep<<<1, SCAN_THREADS, 0, c->stream>>>(d_in, d_out, nullptr, n);
        c->ctr.launches++;
        return YA_OK;
    }
    if (tiles > tmp_words) return ya_fail(c, YA_E_STATE, "scan scratch too small");
    scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(d_in, tmp, n);
    c->ctr.launches++;
    int rc = scan_rec(c, tmp, tmp, tiles, tmp + tiles, tmp_words - tiles);   // in-place scan of tile sums
    if (rc != YA_OK) return rc;
    scan_downsweep<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(d_in, d_out, tmp, n);
    c->ctr.launches++;
    return YA_OK;
}

// d_in and d_out may alias.  If d_total != nullptr it receives the grand total (device memory).
int ya_exclusive_scan_u32(ya_ctx *c, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total)
{
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    size_t words = tiles + tiles / SCAN_TILE + 64 + 2;
    YA_CUDA(c, c->d_scan_tmp.reserve(words * 4 + 16));
    uint32_t *tmp = c->d_scan_tmp.as<uint32_t>();
    uint32_t *last = tmp;            // word 0 keeps in[n-1] when aliasing would destroy it
    if (d_total && n) {
        YA_CUDA(c, cudaMemcpyAsync(last, d_in + (n - 1), 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    int rc = scan_rec(c, d_in, d_out, n, tmp + 2, words - 2);
    if (rc != YA_OK) return rc;
    if (d_total) {
        if (n) {
            scan_write_total<<<1, 1, 0, c->stream>>>(last, d_out, n, d_total);
            c->ctr.launches++;
        } else {
            YA_CUDA(c, cudaMemsetAsync(d_total, 0, 4, c->stream));
        }
    }
    YA_CUDA(c, cudaGetLastError());
    return YA_OK;
}
