// peak.cu -- on-device micro-benchmarks for roofline denominators that MEASURED_PEAKS.json
// does not carry: the sustained INT32 issue rate.
#include "common.cuh"

#define PK_ITERS   4096
#define PK_UNROLL  8

// 8 independent accumulators, 1 IADD3-class op each per inner step.
__global__ void int32_add_kernel(int *out, int seed)
{
    int a0 = seed + threadIdx.x, a1 = a0 ^ 1, a2 = a0 ^ 2, a3 = a0 ^ 3, a4 = a0 ^ 4, a5 = a0 ^ 5, a6 = a0 ^ 6, a7 = a0 ^ 7;
    const int k = seed | 1;
#pragma unroll 4
    for (int i = 0; i < PK_ITERS; i++) {
        a0 += k ^ a1; a1 += k ^ a2; a2 += k ^ a3; a3 += k ^ a4;      // LOP3 + IADD each
        a4 += k ^ a5; a5 += k ^ a6; a6 += k ^ a7; a7 += k ^ a0;
    }
    if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x7fffffff) out[0] = a0;
}

// The DP-cell mix: add, compare, select, max -- 4 independent chains.
__global__ void int32_mix_kernel(int *out, int seed)
{
    int v0 = seed + threadIdx.x, v1 = v0 + 3, v2 = v0 + 5, v3 = v0 + 7;
    int e0 = v0 - 9, e1 = v1 - 9, e2 = v2 - 9, e3 = v3 - 9;
    const int g = (seed & 3) + 2, o = (seed & 7) + 7;
#pragma unroll 4
    for (int i = 0; i < PK_ITERS; i++) {
        // per chain: 2 add, 1 compare+select pair, 1 max, 1 add  (6 ops)
        int c0 = e0 - g, n0 = v0 - o; e0 = (c0 >= n0) ? c0 : n0; v0 = max(v0 + 1, e0) + (e0 > v1 ? 1 : -3);
        int c1 = e1 - g, n1 = v1 - o; e1 = (c1 >= n1) ? c1 : n1; v1 = max(v1 + 1, e1) + (e1 > v2 ? 1 : -3);
        int c2 = e2 - g, n2 = v2 - o; e2 = (c2 >= n2) ? c2 : n2; v2 = max(v2 + 1, e2) + (e2 > v3 ? 1 : -3);
        int c3 = e3 - g, n3 = v3 - o; e3 = (c3 >= n3) ? c3 : n3; v3 = max(v3 + 1, e3) + (e3 > v0 ? 1 : -3);
    }
    if ((v0 ^ v1 ^ v2 ^ v3 ^ e0 ^ e1 ^ e2 ^ e3) == 0x7fffffff) out[0] = v0;
}

extern "C" int ya_measure_int32_peak(ya_ctx *c, double *giops_add, double *giops_mix)
{
    if (!c) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    YA_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    YA_CUDA(c, c->d_misc.reserve(256));
    int *d_out = c->d_misc.as<int>();
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    for (int which = 0; which < 2; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            YA_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
            if (which == 0) int32_add_kernel<<<blocks, threads, 0, c->stream>>>(d_out, rep + 1);
            else            int32_mix_kernel<<<blocks, threads, 0, c->stream>>>(d_out, rep + 1);
            YA_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
            YA_CUDA(c, cudaEventSynchronize(c->ev[1]));
            float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
            if (rep > 0 && ms < best) best = ms;
            c->ctr.launches++;
        }
        // ops per thread: add kernel 8 chains x 2 ops; mix kernel 4 chains x 8 ops (2 sub, cmp, sel, add, max, cmp, sel-add)
        double ops_per_thread = (which == 0) ? (double)PK_ITERS * 16.0 : (double)PK_ITERS * 32.0;
        double g = ops_per_thread * threads * (double)blocks / (best * 1e-3) / 1e9;
        if (which == 0 && giops_add) *giops_add = g;
        if (which == 1 && giops_mix) *giops_mix = g;
    }
    YA_CUDA(c, cudaGetLastError());
    return YA_OK;
}

// Random 4-byte gathers over the resident starting-offset table (4 GiB at K=15): the rate at which
// this GPU's HBM serves independent sector misses, which -- not streaming bandwidth -- bounds the
// seed lookup.  8 independent gathers in flight per thread, all SMs.
__global__ void gather_peak_kernel(const uint32_t *__restrict__ table, uint64_t n_words, uint32_t *out, uint32_t seed, int iters)
{
    uint32_t x = seed ^ (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x = x * 1664525u + 1013904223u;
            uint64_t idx = ((uint64_t)x * n_words) >> 32;
            asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v[k]) : "l"(table + idx));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

extern "C" int ya_measure_gather_peak(ya_ctx *c, double *gather_per_s)
{
    if (!c || !gather_per_s) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    YA_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    YA_CUDA(c, c->d_misc.reserve(256));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 64;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        YA_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
        gather_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->d_so, (uint64_t)c->n_so, c->d_misc.as<uint32_t>(), 77u + rep, iters);
        YA_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
        YA_CUDA(c, cudaEventSynchronize(c->ev[1]));
        float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        if (rep > 0 && ms < best) best = ms;
        c->ctr.launches++;
    }
    *gather_per_s = (double)threads * blocks * iters * 8 / (best * 1e-3);
    YA_CUDA(c, cudaGetLastError());
    return YA_OK;
}
