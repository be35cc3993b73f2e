// index.cu -- build yaha's k-mer index on the device, in the reference's exact layout.
//
// Replaces the 2-pass count/fill of indexFile (Index.c:95-242) by: hash every reference position
// (one thread each, windows containing a non-ACGT code are skipped, Index.c:105-128) ->
// histogram into the starting-offset table -> exclusive scan (Index.c:185-194) -> stable radix
// sort of (hash, position) keys by hash, which leaves each k-mer's list in ascending offset order
// exactly like the reference's second pass (Index.c:201-242).  Only -S 1 (skipDist 1) is built.
// Lists longer than maxHits need the reference's sequential Floyd sampling (Index.c:271-315);
// that case is reported, not approximated.
#include "common.cuh"

__global__ void index_hash_kernel(const uint8_t *__restrict__ bases, uint32_t seq_start, uint32_t n_pos, int K,
                                  uint32_t *__restrict__ counts, uint64_t *__restrict__ keys, uint32_t key_base)
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
    if (bad < 4) {
        atomicAdd(&counts[h], 1u);
        keys[key_base + t] = ((uint64_t)h << 32) | pos;
    } else {
        keys[key_base + t] = ~0ull;                      // sorts behind every real k-mer
    }
}

__global__ void index_take_pos_kernel(const uint64_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ roa)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) roa[i] = (uint32_t)keys[i];
}

__global__ void index_max_count_kernel(const uint32_t *__restrict__ so, size_t n_so, uint32_t *__restrict__ maxc)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    uint32_t m = 0;
    for (; i + 1 < n_so; i += step) { uint32_t c = so[i + 1] - so[i]; if (c > m) m = c; }
    for (int d = 16; d; d >>= 1) { uint32_t o = __shfl_xor_sync(0xffffffffu, m, d); if (o > m) m = o; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(maxc, m);
}


extern "C" ya_ctx *ya_open_build(int device, const ya_params *params, const uint8_t *bases, size_t n_base_bytes,
                                 const uint32_t *seq_start, const uint32_t *seq_len, int n_seq,
                                 uint32_t index_max_hits)
{
    ya_ctx *c = ya_open_common_for_index(device, params);
    if (!c) return nullptr;
    auto fail = [&](const std::string &m) -> ya_ctx * {
        ya_set_open_error(m); ya_close(c); return nullptr;
    };
    if (!bases || n_seq <= 0 || !seq_start || !seq_len) return fail("ya_open_build: bad arguments");
    const int K = params->wordLen;
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
    uint64_t *ka = c->d_keys0.as<uint64_t>(), *kb = c->d_keys1.as<uint64_t>();
    uint32_t kbase = 0;
    for (int s = 0; s < n_seq; s++) {
        if (seq_len[s] < (uint32_t)K) continue;
        uint32_t np = seq_len[s] - K + 1;
        index_hash_kernel<<<(np + 255) / 256, 256, 0, c->stream>>>(c->d_bases, seq_start[s], np, K, c->d_so, ka, kbase);
        kbase += np;
        c->ctr.launches++;
    }
    // counts -> exclusive prefix sums; entry 4^K receives the total (Index.c:185-194)
    uint32_t *d_total = c->d_so + (n_so - 1);
    if (ya_exclusive_scan_u32(c, c->d_so, c->d_so, n_so - 1, d_total) != YA_OK) return fail("index scan failed: " + c->err);
    uint32_t total = 0;
    cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, c->stream);
    if (c->d_misc.reserve(64) != cudaSuccess) return fail("cudaMalloc failed");
    uint32_t *d_max = c->d_misc.as<uint32_t>();
    cudaMemsetAsync(d_max, 0, 4, c->stream);
    index_max_count_kernel<<<148 * 8, 256, 0, c->stream>>>(c->d_so, n_so, d_max);
    uint32_t maxc = 0;
    cudaMemcpyAsync(&maxc, d_max, 4, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("index build: kernel failure");
    if (maxc > index_max_hits)
        return fail("ya_open_build: a k-mer occurs more often than maxHits; its list needs the reference's "
                    "sequential random down-sampling (Index.c:271-315) -- build this index with the host tool");
    if (ya_radix_sort_u64(c, ka, kb, n_keys, 32, 32 + 2 * K) != YA_OK) return fail("index sort failed: " + c->err);
    if ((e = cudaMalloc(&c->d_roa, ((size_t)total + 8) * 4)) != cudaSuccess)
        return fail(std::string("cudaMalloc(roa): ") + cudaGetErrorString(e));
    if (total) index_take_pos_kernel<<<(total + 255) / 256, 256, 0, c->stream>>>(ka, total, c->d_roa);
    cudaMemsetAsync(c->d_roa + total, 0, 32, c->stream);
    c->ctr.launches += 2;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return fail("index build: kernel failure");
    c->n_roa = total;
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
