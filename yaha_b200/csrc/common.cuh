// common.cuh -- internal declarations shared by the translation units of libyaha_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/yaha_b200.h"
#include "finish_reads.h"

#define YA_WORST (-(0x7fffff00))      // "minus infinity" of the reference DP (SW.cpp:356)

// Device-resident growable buffer (never shrinks; reused across batches).
void ya_note_alloc(const char *kind, size_t old_bytes, size_t new_bytes, double t0);   // YA_ALLOC_LOG=1: one stderr line per growth
double ya_now();
// Stream the current thread's device scratch is (re)allocated on: set for the duration of an ABI call
// (AllocScope).  With it DevBuf grows through the stream-ordered allocator -- cudaFree / cudaMalloc in the middle
// of a run wait for every stream of the device and were measured at up to 355 ms under load (YA_ALLOC_LOG).
extern thread_local cudaStream_t ya_tls_alloc_stream;
struct AllocScope {
    cudaStream_t saved;
    explicit AllocScope(cudaStream_t s) : saved(ya_tls_alloc_stream) { ya_tls_alloc_stream = s; }
    ~AllocScope() { ya_tls_alloc_stream = saved; }
};
struct DevBuf {
    void  *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        const double t0 = ya_now(); const size_t old = cap;
        const cudaStream_t s = ya_tls_alloc_stream;
        if (p) { if (s) cudaFreeAsync(p, s); else cudaFree(p); }
        p = nullptr; cap = 0;
        size_t want = 2 * bytes + 256;                // generous: growth is rare and never repeated for the same size class
        cudaError_t e = s ? cudaMallocAsync(&p, want, s) : cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        ya_note_alloc(s ? "device(async)" : "device", old, want, t0);
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// Pinned host staging buffer.
struct PinBuf {
    void  *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        const double t0 = ya_now(); const size_t old = cap;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = 2 * bytes + 256;                // generous: a re-allocation (cudaFree) synchronises the whole device
        if (want < ((size_t)256 << 10)) want = (size_t)256 << 10;     // (and never small: tiny batches of varying size must not re-pin)
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        ya_note_alloc("pinned", old, want, t0);
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// Job descriptor as the DP kernels see it (built on the host from ya_dp_job).
struct DevJob {
    uint64_t tb_off;      // first traceback cell of this job in the scratch (uint16 units)
    uint32_t rOff;        // first reference base of the window (already clamped)
    uint32_t qIdx;        // index of the first query code in the strand's code array
    uint32_t ops_off;     // first slot of this job in the raw op scratch
    uint16_t rLen;        // window length
    uint16_t qLen;        // rows
    uint16_t lb, rb;      // left / right band (band coordinates, SW.cpp:844-871)
    uint8_t  kind;        // YA_DP_*
    uint8_t  strand;
    uint8_t  layout;      // 0: cell (i,j) at i*stride + j ; 1: skewed by lane, (i + j/C)*stride + j
    uint8_t  colsPerLane; // C of the skewed layout
    uint32_t stride;      // traceback row stride in cells
    uint32_t ops_cap;     // slots available in the raw op scratch
    uint32_t rows_off;    // generic kernel: first int of this job's row state scratch
};

struct DevJobOut {
    int32_t  score;       // raw DP score (extension: maxScore; global: last cell)
    int32_t  maxi, maxj;  // traceback start (band coordinates)
    uint32_t n_ops;       // filled by the traceback kernel
    uint32_t cells_lo, cells_hi;   // 64-bit cell count split (avoids alignment padding)
};

// YA_GPU_LOCK=1: one device phase at a time per GPU -- the pipelines of a process take turns for their
// launch..sync sections instead of sharing the SMs (CUDA-event timings of a kernel are then not stretched
// by another pipeline's launches).  Off by default: most calls are bound by launch and serial-chain latency,
// and overlapping them is worth more than clean timings.
std::mutex &ya_device_mutex(int device);
bool ya_device_turns();
// Waits for a stream the way YA_SYNC asks.  Default "nap": poll cudaStreamQuery and sleep ~20 us between
// polls -- the thread driving a context shares the machine with the host's worker threads, so it must not
// burn a core while the device works.  "yield" polls with sched_yield; "spin" / "block" are the driver's own
// cudaStreamSynchronize (with cudaDeviceScheduleBlockingSync for "block").
cudaError_t ya_stream_wait(cudaStream_t st);
cudaError_t ya_event_wait(cudaEvent_t ev);
std::mutex &ya_bulk_mutex(int device);

struct DeviceTurn {
    std::unique_lock<std::mutex> lk;
    explicit DeviceTurn(int device) : lk(ya_device_mutex(device), std::defer_lock) { if (ya_device_turns()) lk.lock(); }
    void done() { if (lk.owns_lock()) lk.unlock(); }
};

struct ya_ctx {
    int          device = -1;
    ya_params    P{};
    cudaStream_t own_stream = nullptr;    // = prio_streams[prio_level]
    cudaStream_t prio_streams[4] = {nullptr, nullptr, nullptr, nullptr};   // ya_set_priority: level 0 = the device's greatest priority
    int          prio_level = 0;
    cudaStream_t bulk_stream = nullptr;   // lowest priority: DP calls with many jobs
    cudaStream_t stream = nullptr;
    std::string  err;
    // index + genome (device resident)
    uint32_t *d_so = nullptr;   size_t n_so = 0;
    uint32_t *d_roa = nullptr;  size_t n_roa = 0;
    uint8_t  *d_bases = nullptr; size_t n_base_bytes = 0;
    uint32_t  maxROff = 0;
    bool      owns_index = true;
    bool      peer_direct = false;     // ya_open_peer: the replica was copied with peer access enabled (NVLink, no host staging)
    uint32_t *d_lowmask = nullptr;   // 2^20-bit filter: k-mers (hash & 0xFFFFF) occurring in the first 32 K reference bases
    // uploaded read batch
    int       n_reads = 0;
    uint64_t  total_bases = 0;
    DevBuf    d_codes_fwd, d_codes_rev, d_read_off;
    std::vector<uint64_t> h_read_off;
    // seed stage scratch
    DevBuf    d_seg_probe_off, d_cnt, d_soff, d_hit_off, d_keys0, d_keys1, d_scan_tmp, d_hist;
    DevBuf    d_fragflag, d_fragidx, d_frags_all, d_frag_seg, d_regflag, d_regidx, d_regstart,
              d_keep, d_keepidx, d_frags_out, d_region_out, d_strand_out, d_misc;
    DevBuf    d_fc_count, d_fc_work, d_fc_tmp, d_fc_path, d_fc_nodes, d_fc_used, d_fc_clumps, d_fc_slot;   // ya_form_clumps
    DevBuf    d_pc_path, d_pc_gaps, d_pc_prep, d_pc_jobs;                                       // ya_prepare_clumps
    // ya_align_batch: the reads as text, per-clump assembly records, per-read output plan, SAM text
    DevBuf    d_chars, d_quals, d_ids, d_fin, d_asm_recs, d_asm_ops, d_fr_outs, d_text, d_out_tab;
    ya_out_params out{}; fr_params fr{}; bool out_set = false;
    size_t    text_pending = 0;         // text left on the device by a ya_align_batch that returned YA_E_CAPACITY
    std::vector<float> ext_iv;          // [start, end) of every bulk extension launch, ms since the device's base event
    bool      dpr_ran = false;          // the last device DP round launched kernels (its events are valid)
    DevBuf    d_dpr;                                                                           // plan of a device-born DP round
    bool      dpr_bulk_packed = false; int dpr_packed_launches = 0;
    bool      big_sort_attr = false;    // seg_sort_kernel<8192,256> has its 64 KB shared-memory opt-in on this context's device
    bool      fc_valid = false;         // the clumps of the last ya_form_clumps call are on the device
    int       seed_chunks = 0;          // chunks of the last ya_seed_frags call (its survivors stay on the device when 1)
    size_t    seed_nkeep = 0;           // survivors of that call
    PinBuf    h_stage, h_stage2, h_stage3;
    // dp stage scratch
    DevBuf    d_jobs, d_jobout, d_tb, d_rows, d_ops_raw, d_ops_cnt, d_ops_off, d_ops_out, d_res;
    PinBuf    h_jobs, h_res, h_ops;
    std::vector<std::vector<uint32_t>> sw_lists;      // ya_sw_batch host scratch (reused)
    std::vector<uint32_t> sw_live_of, sw_flat, sw_cnt, sw_tmp;
    size_t ops_pending = 0;          // ops left on the device by a ya_sw_batch that returned YA_E_CAPACITY
    std::vector<uint32_t> seed_small, seed_big, seed_koff;   // ya_seed_frags host scratch (segment id lists, key offsets)
    // timing
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ya_counters ctr{};
};

#define YA_CUDA(ctx, call)                                                                  \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
            return YA_E_CUDA;                                                               \
        }                                                                                   \
    } while (0)

static inline int ya_fail(ya_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    return code;
}

// scan.cu
int ya_exclusive_scan_u32(ya_ctx *c, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total);
// seed.cu: stable LSD radix sort of 64-bit keys on bits [lo_bit, hi_bit); result ends up in `a`
int ya_radix_sort_u64(ya_ctx *c, uint64_t *&a, uint64_t *&b, uint32_t n, int lo_bit, int hi_bit);
// seed.cu / sw.cu / peak.cu / index.cu hold the C-ABI entry points declared in yaha_b200.h

// ctx.cu
cudaEvent_t ya_device_base_event(int device, cudaStream_t st);   // recorded once per device and process: a common time origin
void ya_note_ext_interval(ya_ctx *c, cudaEvent_t a, cudaEvent_t b);
int ya_build_lowmask(ya_ctx *c, const uint8_t *host_bases, size_t n_base_bytes);
ya_ctx *ya_open_common_for_index(int device, const ya_params *params);
void ya_set_open_error(const std::string &m);
