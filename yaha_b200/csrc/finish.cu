// finish.cu -- SURVEY.md section 8f, rows N2 and N4: the whole per-read path of a batch in one call (ya_align_batch).
//
//   encode_reads_kernel   characters -> 4-bit codes of both strands (Query.c:161-168, Math.c:141-157)
//   [seed.cu]             seed lookup, hits -> fragments -> regions, survivors kept on the device
//   [clumps.cu]           fragment graph -> clumps of seed fragments; phase 1 of alignClump, DP jobs born on the device
//   [sw.cu]               ya_sw_device_round: job layout, fill kernels, traceback -- answers stay on the device
//   assemble_kernel       csrc/assemble_clumps.h per clump: splice the answers, both extensions, scoreClump's verdict
//                         (AlignHelpers.c:251-366, AlignExtFrag.cpp:64-143)
//   finish_kernel         csrc/finish_reads.h per read: OQC / filter by similarity / mapping quality
//                         (GraphPath.cpp:294-1086), which records to print and how long their SAM text is
//   format_kernel         the SAM text itself (AlignOutput.c:115-289), every read at its place in read order
//
// A read whose control flow leaves the straight path (a clump that must be split, a strand the fragment-graph kernel left
// out, too many clumps) gets status 1 and no text: the caller runs it through the call-by-call ABI.  Per batch the host
// synchronises a handful of times for sizes (hit count, survivor count, job count, scratch totals, text length); no per-read
// or per-job data crosses PCIe except the reads going up and the text coming down.
#include "common.cuh"
#include "finish_reads.h"
#include <math.h>
#include <algorithm>

int ya_seed_frags_impl(ya_ctx *c, ya_frag_batch *out, bool device_only);
int ya_form_clumps_impl(ya_ctx *c, ya_clump_batch *out, bool device_only);
int ya_prepare_clumps_impl(ya_ctx *c, ya_prep_batch *out, bool device_only, size_t *n_ext);
int ya_sw_device_round(ya_ctx *c, uint32_t n_jobs, uint32_t n_ext, unsigned long long *d_acct, size_t *raw_slots);

__constant__ uint8_t c_code_of_char[256];
__constant__ uint8_t c_comp_code[16] = {2, 3, 0, 1, 4, 12, 7, 6, 9, 8, 15, 11, 5, 13, 14, 10};

// one warp per read: forward codes in place order, reverse-complement codes mirrored (both accesses coalesced)
__global__ void encode_reads_kernel(const char *__restrict__ chars, const uint64_t *__restrict__ off, int n_reads,
                                    uint8_t *__restrict__ fwd, uint8_t *__restrict__ rev)
{
    const int lane = threadIdx.x & 31;
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n_reads; r += warps) {
        const uint64_t s = off[r], e = off[r + 1];
        for (uint64_t i = s + lane; i < e; i += 32) {
            const uint8_t cd = c_code_of_char[(uint8_t)chars[i]];
            fwd[i] = cd;
            rev[s + (e - 1 - i)] = c_comp_code[cd];
        }
    }
}

struct AsmDev {
    const ya_clump_rec *clumps; const uint32_t *count, *first;
    const ya_frag *path; const ya_gap_rec *gaps; const ya_prep_rec *prep;
    const ya_dp_result *res; const ya_op *rops;
};

// one thread per clump (slot_strand, written by form_clumps_kernel, says which slots of the clump array are clumps and whose)
__global__ void __launch_bounds__(128)
assemble_kernel(uint32_t n_slots, const uint32_t *__restrict__ slot_strand, const uint64_t *__restrict__ read_off, AsmDev D,
                const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd, const uint8_t *__restrict__ rev, ac_params P,
                ya_asm_rec *__restrict__ recs, ya_op *__restrict__ out_ops, uint32_t out_cap, uint32_t *__restrict__ out_used,
                unsigned long long *__restrict__ acct)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const uint32_t sp1 = slot_strand[i];
    if (sp1 == 0) return;
    const uint32_t s = sp1 - 1, r = s >> 1;
    const uint64_t base = read_off[r];
    const int readLen = (int)(read_off[r + 1] - base);
    const uint8_t *q = ((s & 1) ? rev : fwd) + base;
    const ya_clump_rec cr = D.clumps[i];
    const ya_prep_rec pr = D.prep[i];
    const ya_gap_rec *g = D.gaps + pr.gap_first;
    const uint32_t bound = ac_ops_bound((int)cr.n, g, pr.n_gaps, &pr, D.res, D.rops);
    const uint32_t at = atomicAdd(out_used, bound);
    ya_asm_rec rec;
    if ((uint64_t)at + bound > out_cap) {                                  // (cannot happen: the caller sized the array by the same bound)
        atomicOr(&acct[2], 8ull);
        memset(&rec, 0, sizeof rec); rec.verdict = YA_ASM_SPLIT;
    } else {
        if (ac_assemble_clump(&P, bases, q, readLen, D.path + cr.first, (int)cr.n, g, pr.n_gaps, &pr, D.res, D.rops, out_ops + at, &rec) != 0) {
            atomicOr(&acct[2], 16ull);                                     // extension plan diverged (fatal on the host as well)
            rec.verdict = YA_ASM_SPLIT;
        }
        rec.ops_off = at;
    }
    recs[i] = rec;
}

struct ReadText { const char *chars, *quals, *ids; const uint32_t *id_off; };

// one thread per read: which records to print (fr_finish_read) and the length of their text
__global__ void __launch_bounds__(64)
finish_kernel(int n_reads, const uint64_t *__restrict__ read_off, const ya_strand_frags *__restrict__ strands,
              const uint32_t *__restrict__ count, const uint32_t *__restrict__ first,
              const ya_asm_rec *__restrict__ recs, const ya_op *__restrict__ asm_ops, const uint8_t *__restrict__ bases,
              const uint8_t *__restrict__ fwd, const uint8_t *__restrict__ rev, ReadText T, fr_params P,
              fr_out *__restrict__ outs, uint32_t *__restrict__ n_outs, uint32_t *__restrict__ primary_count,
              uint32_t *__restrict__ text_len, uint8_t *__restrict__ status)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint64_t base = read_off[r];
    const int readLen = (int)(read_off[r + 1] - base);
    fr_clump cl[FR_MAX_NODES];
    uint32_t asmIdx[FR_MAX_NODES];
    int n = 0;
    bool handBack = false;
    for (int st = 0; st < 2 && !handBack; st++) {
        const int s = 2 * r + st;
        const uint32_t nc = count[s], c0 = first[s];
        if (nc == 0xFFFFFFFFu) { handBack = true; break; }
        for (uint32_t k = 0; k < nc; k++) {
            const ya_asm_rec *a = &recs[c0 + k];
            if (a->verdict == YA_ASM_SPLIT) { handBack = true; break; }
            if (a->verdict != YA_ASM_SCORED) continue;
            if (n == FR_MAX_NODES) { handBack = true; break; }
            cl[n].rec = a; cl[n].ops = asm_ops + a->ops_off; cl[n].reversed = st;
            asmIdx[n] = c0 + k;
            n++;
        }
    }
    uint32_t len = 0, nOut = 0;
    int primaries = 0;
    if (!handBack && n > 0) {
        fr_node g[FR_MAX_NODES];
        fr_out o[FR_MAX_NODES];
        const int k = fr_finish_read(&P, fwd + base, readLen, cl, n, g, o, &primaries);
        if (k < 0) handBack = true;
        else {
            // the read owns the slots of its two strands' surviving fragments (at least as many as its clumps); an empty
            // strand has no place of its own, so the run starts at the first strand that has fragments
            fr_out *dst = outs + (strands[2 * r].n_frags ? strands[2 * r].first : strands[2 * r + 1].first);
            const uint32_t idLen = T.id_off[r + 1] - T.id_off[r];
            for (int q = 0; q < k; q++) {
                len += (uint32_t)fr_format_record(&P, bases, T.ids + T.id_off[r], (int)idLen, T.chars + base, T.quals ? T.quals + base : nullptr,
                                                  rev + base, readLen, &cl[o[q].clump], &o[q], primaries, nullptr);
                fr_out w = o[q];
                w.clump = (int32_t)asmIdx[o[q].clump];                     // absolute record index for the format kernel
                dst[q] = w;
            }
            nOut = (uint32_t)k;
        }
    }
    if (handBack) { len = 0; nOut = 0; }
    status[r] = handBack ? 1 : 0;
    n_outs[r] = nOut; primary_count[r] = (uint32_t)primaries; text_len[r] = len;
}

// one WARP per read writes its records at the read's place in the batch's text (finish_reads.h: all lanes make the same
// calls; single characters come from lane 0, the long runs are strided over the lanes)
__global__ void __launch_bounds__(128)
format_kernel(int n_reads, const uint64_t *__restrict__ read_off, const ya_strand_frags *__restrict__ strands,
              const ya_asm_rec *__restrict__ recs, const ya_op *__restrict__ asm_ops, const uint8_t *__restrict__ bases,
              const uint8_t *__restrict__ rev, ReadText T, fr_params P, const fr_out *__restrict__ outs,
              const uint32_t *__restrict__ n_outs, const uint32_t *__restrict__ primary_count, const uint32_t *__restrict__ text_off,
              char *__restrict__ text)
{
    const int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= n_reads) return;
    const uint32_t k = n_outs[r];
    if (k == 0) return;
    const uint64_t base = read_off[r];
    const int readLen = (int)(read_off[r + 1] - base);
    const fr_out *o = outs + (strands[2 * r].n_frags ? strands[2 * r].first : strands[2 * r + 1].first);
    const uint32_t idLen = T.id_off[r + 1] - T.id_off[r];
    char *w = text + text_off[r];
    for (uint32_t q = 0; q < k; q++) {
        fr_clump c;
        c.rec = &recs[o[q].clump]; c.ops = asm_ops + c.rec->ops_off; c.reversed = (o[q].status & FR_REVERSED) != 0;
        w += fr_format_record(&P, bases, T.ids + T.id_off[r], (int)idLen, T.chars + base, T.quals ? T.quals + base : nullptr, rev + base, readLen,
                              &c, &o[q], (int)primary_count[r], w);
    }
}

// exclusive scan of the per-read text lengths by one block (n_reads is a few thousand .. tens of thousands); also widens
// the offsets to 64 bits for the caller and counts the reads handed back
__global__ void __launch_bounds__(1024)
text_scan_kernel(const uint32_t *__restrict__ len, const uint8_t *__restrict__ status, int n, uint32_t *__restrict__ off32,
                 uint64_t *__restrict__ off64, unsigned long long *__restrict__ totals)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    __shared__ unsigned int handed;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry = 0; handed = 0; }
    __syncthreads();
    unsigned int myHanded = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = (i < n) ? len[i] : 0ull;
        if (i < n && status[i]) myHanded++;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            unsigned long long x = wsum[lane], xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += t; }
            wsum[lane] = xi - x;
        }
        __syncthreads();
        const unsigned long long ex = carry + wsum[w] + inc - v;
        if (i < n) { off32[i] = (uint32_t)ex; off64[i] = ex; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = ex + v;
        __syncthreads();
    }
    if (myHanded) atomicAdd(&handed, myHanded);
    __syncthreads();
    if (threadIdx.x == 0) { off64[n] = carry; totals[0] = carry; totals[1] = handed; }
}

// -----------------------------------------------------------------------------------------------------------------
extern "C" int ya_set_output(ya_ctx *c, const ya_out_params *o, int n_seq, const char *const *seq_names,
                             const uint32_t *seq_start, const uint32_t *seq_len)
{
    if (!c || !o || n_seq <= 0 || !seq_names || !seq_start || !seq_len) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    c->out = *o;
    // characters -> codes: T C A G N B D H K M R S V W X Y (Math.c:154), upper and lower case, U as T, everything else X
    uint8_t tab[256];
    memset(tab, 14, sizeof tab);
    static const char kChar[16] = {'T', 'C', 'A', 'G', 'N', 'B', 'D', 'H', 'K', 'M', 'R', 'S', 'V', 'W', 'X', 'Y'};
    for (int i = 0; i < 16; i++) { tab[(int)kChar[i]] = (uint8_t)i; tab[(int)kChar[i] + 32] = (uint8_t)i; }
    tab['U'] = tab['u'] = 0;
    YA_CUDA(c, cudaMemcpyToSymbol(c_code_of_char, tab, 256));
    // sequence table
    std::vector<uint32_t> nameOff((size_t)n_seq + 1);
    std::string names;
    for (int i = 0; i < n_seq; i++) { nameOff[(size_t)i] = (uint32_t)names.size(); names += seq_names[i]; }
    nameOff[(size_t)n_seq] = (uint32_t)names.size();
    // break-point penalty steps (GraphPath.cpp:1018-1020) with this process' log10 -- the C library the reference binary uses
    auto bpp = [&](uint32_t d) {
        double lg = log10((double)d);
        if (lg > o->maxBPLog) lg = (double)o->maxBPLog;
        return (int)(lg * o->BPCost + 0.5);
    };
    std::vector<uint32_t> steps;
    const int b0 = bpp(11), b1 = bpp(0xFFFFFFFFu);
    if (b1 - b0 > 4096 || b1 < b0) return ya_fail(c, YA_E_ARG, "ya_set_output: unsupported -BP / -MGDP combination");
    for (int b = b0 + 1; b <= b1; b++) {                                 // smallest distance whose penalty reaches b
        uint64_t lo = 11, hi = 0xFFFFFFFFull;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (bpp((uint32_t)mid) >= b) hi = mid; else lo = mid + 1; }
        steps.push_back((uint32_t)lo);
    }
    const size_t bytes = ((size_t)3 * n_seq + 1 + steps.size() + 16) * 4 + names.size() + 64;
    YA_CUDA(c, c->d_out_tab.reserve(bytes));
    uint32_t *d_start = c->d_out_tab.as<uint32_t>(), *d_len = d_start + n_seq, *d_noff = d_len + n_seq, *d_steps = d_noff + n_seq + 1;
    char *d_names = (char *)(d_steps + steps.size() + 1);
    YA_CUDA(c, cudaMemcpy(d_start, seq_start, (size_t)n_seq * 4, cudaMemcpyHostToDevice));
    YA_CUDA(c, cudaMemcpy(d_len, seq_len, (size_t)n_seq * 4, cudaMemcpyHostToDevice));
    YA_CUDA(c, cudaMemcpy(d_noff, nameOff.data(), ((size_t)n_seq + 1) * 4, cudaMemcpyHostToDevice));
    if (!steps.empty()) YA_CUDA(c, cudaMemcpy(d_steps, steps.data(), steps.size() * 4, cudaMemcpyHostToDevice));
    if (!names.empty()) YA_CUDA(c, cudaMemcpy(d_names, names.data(), names.size(), cudaMemcpyHostToDevice));
    fr_params &F = c->fr;
    F.GOCost = c->P.GOCost; F.GECost = c->P.GECost; F.RCost = c->P.RCost; F.MScore = c->P.MScore;
    F.OQC = o->OQC; F.FBS = o->FBS; F.OQCMinNonOverlap = o->OQCMinNonOverlap; F.BPCost = o->BPCost; F.maxBPLog = o->maxBPLog;
    F.FBS_PSLength = o->FBS_PSLength; F.FBS_PSScore = o->FBS_PSScore;
    F.hardClip = o->hardClip; F.fastq = o->fastq;
    F.n_seq = n_seq; F.seq_start = d_start; F.seq_len = d_len; F.seq_name_off = d_noff; F.seq_names = d_names;
    F.bpp_base = b0; F.n_bpp = (int)steps.size(); F.bpp_dist = d_steps;
    c->out_set = true;
    return YA_OK;
}

extern "C" int ya_align_fetch_text(ya_ctx *c, char *text, size_t text_cap)
{
    if (!c) return YA_E_ARG;
    if (c->text_pending == 0) return ya_fail(c, YA_E_STATE, "ya_align_fetch_text: no text pending");
    if (!text || text_cap < c->text_pending) return ya_fail(c, YA_E_CAPACITY, "text buffer too small");
    YA_CUDA(c, cudaSetDevice(c->device));
    YA_CUDA(c, cudaMemcpyAsync(text, c->d_text.p, c->text_pending, cudaMemcpyDeviceToHost, c->stream));
    YA_CUDA(c, ya_stream_wait(c->stream));
    return YA_OK;
}

// YA_PROF=1: wall time of the phases of ya_align_batch, summed over the process, printed at exit
static double g_prof_ab[8]; static long g_prof_ab_n;
struct AbProfPrinter { ~AbProfPrinter() { if (getenv("YA_PROF") && g_prof_ab_n) fprintf(stderr, "ya_align_batch wall ms per call (%ld calls): upload+encode enqueue %.3f | seed stage %.3f | "
    "clumps+prepare %.3f | dp plan+launch %.3f | assemble+finish (sync) %.3f | format+d2h (sync) %.3f\n", g_prof_ab_n, 1e3 * g_prof_ab[0] / g_prof_ab_n,
    1e3 * g_prof_ab[1] / g_prof_ab_n, 1e3 * g_prof_ab[2] / g_prof_ab_n, 1e3 * g_prof_ab[3] / g_prof_ab_n, 1e3 * g_prof_ab[4] / g_prof_ab_n, 1e3 * g_prof_ab[5] / g_prof_ab_n); } } g_ab_prof_printer;

extern "C" int ya_align_batch(ya_ctx *c, ya_text_batch *b)
{
    double tq = ya_now();
    auto lap = [&](int k) { const double t = ya_now(); g_prof_ab[k] += t - tq; tq = t; };
    if (!c || !b || b->n_reads < 0 || !b->text_off || !b->status || (b->n_reads && (!b->chars || !b->offsets || !b->ids || !b->id_off)))
        return YA_E_ARG;
    if (!c->out_set) return ya_fail(c, YA_E_STATE, "ya_align_batch: call ya_set_output first");
    if (c->out.fastq && b->n_reads && !b->quals) return ya_fail(c, YA_E_ARG, "ya_align_batch: FASTQ output needs the quality characters");
    YA_CUDA(c, cudaSetDevice(c->device));
    b->text_len = 0; b->text_needed = 0; b->n_handed_back = 0;
    c->text_pending = 0;
    const int n = b->n_reads;
    b->text_off[0] = 0;
    if (n == 0) { c->n_reads = 0; return YA_OK; }
    if (n > (1 << 16)) return ya_fail(c, YA_E_STATE, "ya_align_batch: at most 65536 reads per call");
    cudaStream_t st = c->stream;
    AllocScope allocScope(st);
    // ---- reads up, codes of both strands made on the device
    c->n_reads = n;
    c->h_read_off.assign(b->offsets, b->offsets + n + 1);
    if (c->h_read_off[0] != 0) return ya_fail(c, YA_E_ARG, "offsets[0] must be 0");
    for (int r = 0; r < n; r++) {
        if (c->h_read_off[(size_t)r + 1] < c->h_read_off[(size_t)r] || c->h_read_off[(size_t)r + 1] - c->h_read_off[(size_t)r] > 32767)
            return ya_fail(c, YA_E_ARG, "read length must be in 0..32767 (16-bit query offsets, Math.h:104)");
        if (b->id_off[r + 1] < b->id_off[r]) return ya_fail(c, YA_E_ARG, "id offsets must ascend");
    }
    const uint64_t total = c->h_read_off[(size_t)n];
    if (total >= 0xFFFF0000ull) return ya_fail(c, YA_E_ARG, "at most 2^32 - 65536 bases per batch (32-bit code offsets): split the batch");
    c->total_bases = total;
    const size_t idBytes = b->id_off[n];
    YA_CUDA(c, c->d_codes_fwd.reserve(total + 64));
    YA_CUDA(c, c->d_codes_rev.reserve(total + 64));
    YA_CUDA(c, c->d_read_off.reserve((size_t)(n + 1) * 8));
    YA_CUDA(c, c->d_chars.reserve(total + 64));
    if (b->quals) YA_CUDA(c, c->d_quals.reserve(total + 64));
    YA_CUDA(c, c->d_ids.reserve(idBytes + (size_t)(n + 1) * 4 + 64));
    uint32_t *d_id_off = c->d_ids.as<uint32_t>();
    char *d_ids = (char *)(d_id_off + n + 1);
    YA_CUDA(c, cudaMemcpyAsync(c->d_read_off.p, b->offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    YA_CUDA(c, cudaMemcpyAsync(d_id_off, b->id_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    if (idBytes) YA_CUDA(c, cudaMemcpyAsync(d_ids, b->ids, idBytes, cudaMemcpyHostToDevice, st));
    if (total) {
        YA_CUDA(c, cudaMemcpyAsync(c->d_chars.p, b->chars, total, cudaMemcpyHostToDevice, st));
        if (b->quals) YA_CUDA(c, cudaMemcpyAsync(c->d_quals.p, b->quals, total, cudaMemcpyHostToDevice, st));
        int blocks = (n + 7) / 8; if (blocks > 148 * 16) blocks = 148 * 16;
        encode_reads_kernel<<<blocks, 256, 0, st>>>(c->d_chars.as<char>(), c->d_read_off.as<uint64_t>(), n, c->d_codes_fwd.as<uint8_t>(),
                                                   c->d_codes_rev.as<uint8_t>());
        c->ctr.launches++;
    }
    // ---- stages 1 + 2, fragment graph, phase 1 of the alignment: all left on the device
    lap(0);
    ya_frag_batch fb; memset(&fb, 0, sizeof fb);
    int rc = ya_seed_frags_impl(c, &fb, true);
    if (rc != YA_OK) return rc;
    lap(1);
    const size_t nk = c->seed_nkeep;
    const int n_seg = 2 * n;
    YA_CUDA(c, c->d_fin.reserve(4 * 8 + (size_t)n * (4 * 4 + 8 + 1) + 64 + 8));
    unsigned long long *d_acct = c->d_fin.as<unsigned long long>();            // [0] cells [1] packed cells [2] error flags [3] spare
    unsigned long long *d_text_tot = d_acct + 4;                                // text bytes, reads handed back
    uint64_t *d_text_off64 = (uint64_t *)(d_text_tot + 2);                      // n + 1
    uint32_t *d_nouts = (uint32_t *)(d_text_off64 + n + 1), *d_prim = d_nouts + n, *d_tlen = d_prim + n, *d_toff = d_tlen + n;
    uint8_t *d_status = (uint8_t *)(d_toff + n);
    YA_CUDA(c, cudaMemsetAsync(d_acct, 0, 6 * 8, st));
    size_t rawSlots = 0;
    ya_prep_batch pb; memset(&pb, 0, sizeof pb);
    size_t nExt = 0;
    if (nk) {
        ya_clump_batch cb; memset(&cb, 0, sizeof cb);
        cb.maxDesert = c->out.maxDesert; cb.minNonOverlap = c->out.minNonOverlap;
        rc = ya_form_clumps_impl(c, &cb, true);
        if (rc != YA_OK) return rc;
        rc = ya_prepare_clumps_impl(c, &pb, true, &nExt);
        if (rc != YA_OK) return rc;
        lap(2);
        // ---- the first DP round, answers on the device
        rc = ya_sw_device_round(c, (uint32_t)pb.n_jobs, (uint32_t)nExt, d_acct, &rawSlots);
        if (rc != YA_OK) return rc;
    }
    uint32_t *d_count = c->d_fc_count.as<uint32_t>(), *d_first = d_count + n_seg;
    if (!nk) {                                                                   // no survivor at all: every strand has no clump
        YA_CUDA(c, c->d_fc_count.reserve((size_t)n_seg * 8 + 64));
        d_count = c->d_fc_count.as<uint32_t>(); d_first = d_count + n_seg;
        YA_CUDA(c, cudaMemsetAsync(d_count, 0, (size_t)n_seg * 8, st));
    }
    lap(3);
    // ---- splice + score every clump, then finish every read
    const size_t asmCap = rawSlots + 2 * nk + 64;                                // runs: DP answers + one per seed piece + one per closed-form gap
    if (asmCap >= 0xFFFF0000ull) return ya_fail(c, YA_E_STATE, "ya_align_batch: batch too large for 32-bit run offsets");
    YA_CUDA(c, c->d_asm_recs.reserve((nk + 1) * sizeof(ya_asm_rec)));
    YA_CUDA(c, c->d_asm_ops.reserve(asmCap * sizeof(ya_op) + 64));
    YA_CUDA(c, c->d_fr_outs.reserve((nk + 1) * sizeof(fr_out)));
    uint32_t *d_asm_used = (uint32_t *)(d_acct + 3);
    if (nk) {
        AsmDev D;
        D.clumps = c->d_fc_clumps.as<ya_clump_rec>(); D.count = d_count; D.first = d_first;
        D.path = c->d_pc_path.as<ya_frag>(); D.gaps = c->d_pc_gaps.as<ya_gap_rec>(); D.prep = c->d_pc_prep.as<ya_prep_rec>();
        D.res = c->d_res.as<ya_dp_result>(); D.rops = c->d_ops_raw.as<ya_op>();
        ac_params AP;
        AP.GOCost = c->P.GOCost; AP.GECost = c->P.GECost; AP.RCost = c->P.RCost; AP.MScore = c->P.MScore;
        AP.minExtLength = c->P.minExtLength; AP.minRawScore = c->out.minRawScore; AP.maxROff = c->maxROff; AP.minIdentity = c->out.minIdentity;
        assemble_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, st>>>((uint32_t)nk, c->d_fc_slot.as<uint32_t>(), c->d_read_off.as<uint64_t>(), D, c->d_bases, c->d_codes_fwd.as<uint8_t>(),
            c->d_codes_rev.as<uint8_t>(), AP, c->d_asm_recs.as<ya_asm_rec>(), c->d_asm_ops.as<ya_op>(), (uint32_t)asmCap, d_asm_used, d_acct);
        c->ctr.launches++;
    }
    ReadText T;
    T.chars = c->d_chars.as<char>(); T.quals = b->quals ? c->d_quals.as<char>() : nullptr; T.ids = d_ids; T.id_off = d_id_off;
    finish_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, c->d_read_off.as<uint64_t>(), c->d_strand_out.as<ya_strand_frags>(), d_count, d_first, c->d_asm_recs.as<ya_asm_rec>(),
        c->d_asm_ops.as<ya_op>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), T, c->fr,
        c->d_fr_outs.as<fr_out>(), d_nouts, d_prim, d_tlen, d_status);
    text_scan_kernel<<<1, 1024, 0, st>>>(d_tlen, d_status, n, d_toff, d_text_off64, d_text_tot);
    c->ctr.launches += 2;
    YA_CUDA(c, c->h_stage3.reserve(1024));
    unsigned long long *h_tot = c->h_stage3.as<unsigned long long>();           // [0..3] acct, [4] text bytes, [5] handed back
    YA_CUDA(c, cudaMemcpyAsync(h_tot, d_acct, 6 * 8, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, ya_stream_wait(st));
    YA_CUDA(c, cudaGetLastError());
    if (h_tot[2] & 0xFFFFFFFFull) {
        char msg[96]; snprintf(msg, sizeof msg, "ya_align_batch: internal error on the device (flags %llx)", h_tot[2] & 0xFFFFFFFFull);
        return ya_fail(c, YA_E_INTERNAL, msg);
    }
    lap(4);
    const size_t textBytes = (size_t)h_tot[4];
    if (textBytes >= 0xFFFF0000ull) return ya_fail(c, YA_E_STATE, "ya_align_batch: more than 4 GB of text, use smaller batches");
    // ---- the text
    YA_CUDA(c, c->d_text.reserve(textBytes + 64));
    if (textBytes) {
        format_kernel<<<(n + 3) / 4, 128, 0, st>>>(n, c->d_read_off.as<uint64_t>(), c->d_strand_out.as<ya_strand_frags>(), c->d_asm_recs.as<ya_asm_rec>(),
            c->d_asm_ops.as<ya_op>(), c->d_bases, c->d_codes_rev.as<uint8_t>(), T, c->fr, c->d_fr_outs.as<fr_out>(), d_nouts, d_prim,
            d_toff, c->d_text.as<char>());
        c->ctr.launches++;
    }
    YA_CUDA(c, cudaEventRecord(c->ev[5], st));
    YA_CUDA(c, cudaMemcpyAsync(b->text_off, d_text_off64, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, cudaMemcpyAsync(b->status, d_status, (size_t)n, cudaMemcpyDeviceToHost, st));
    const bool fits = textBytes <= b->text_cap && (textBytes == 0 || b->text != nullptr);
    if (fits && textBytes) YA_CUDA(c, cudaMemcpyAsync(b->text, c->d_text.p, textBytes, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, ya_stream_wait(st));
    YA_CUDA(c, cudaGetLastError());
    lap(5); g_prof_ab_n++;
    // counters (the DP kernels' events were recorded by ya_sw_device_round on this stream)
    c->ctr.dp_cells += h_tot[0];
    if (c->dpr_ran) {
        float ms0 = 0, ms1 = 0, ms2 = 0;
        if (cudaEventElapsedTime(&ms0, c->ev[0], c->ev[1]) == cudaSuccess) c->ctr.ms_dp += ms0;
        if (cudaEventElapsedTime(&ms1, c->ev[1], c->ev[2]) == cudaSuccess) c->ctr.ms_traceback += ms1;
        if (c->dpr_bulk_packed && cudaEventElapsedTime(&ms2, c->ev[3], c->ev[4]) == cudaSuccess) {
            c->ctr.ms_ext += ms2; c->ctr.ext_cells += h_tot[1]; c->ctr.ext_launches += (uint64_t)c->dpr_packed_launches;
            ya_note_ext_interval(c, c->ev[3], c->ev[4]);
        }
        float ms3 = 0;
        if (cudaEventElapsedTime(&ms3, c->ev[2], c->ev[5]) == cudaSuccess) c->ctr.ms_finish += ms3;
        cudaGetLastError();
    }
    b->text_len = textBytes; b->text_needed = textBytes; b->n_handed_back = (int32_t)h_tot[5];
    c->ctr.reads_handed_back += h_tot[5]; c->ctr.reads_finished += (uint64_t)n - h_tot[5]; c->ctr.text_bytes += textBytes;
    if (!fits) { c->text_pending = textBytes; return ya_fail(c, YA_E_CAPACITY, "text buffer too small"); }
    return YA_OK;
}
