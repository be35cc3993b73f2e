// clumps.cu -- SURVEY.md section 8f, row N1: fragments -> clumps of seed fragments on the device.
//
// ya_form_clumps runs yaha_b200/csrc/form_clumps.h (the region loop of processFragmentsGapped, the fragment
// graph of buildBestClumpFromFragmentRange, insertFragment's overlap chops, cleanUpClump, eliminateFragments:
// QueryMatch.c:170-303, GraphPath.cpp:161-292, AlignHelpers.c:48-193) with one WARP per strand on the
// survivors ya_seed_frags left on the device: lane 0 runs the serial parts (a dozen fragments per strand as a
// rule), the inner loop of the O(n^2) chain DP is strided over the lanes.  Strands with more than
// kMaxStrandFrags fragments (repeat-rich loci) are left to the host: their clump count comes back as
// 0xFFFFFFFF.  ya_prepare_clumps (prepare_clumps.h) then runs one thread per CLUMP.
#define FC_WARP_COOP 1          // form_clumps.h: one warp per strand (must precede every inclusion of the header)
#include "common.cuh"
#include "form_clumps.h"
#include "prepare_clumps.h"

static const uint32_t kMaxStrandFrags = 1024;

__global__ void __launch_bounds__(128)
form_clumps_kernel(const ya_strand_frags *__restrict__ strands, int n_seg, const uint64_t *__restrict__ read_off,
                   const ya_frag *__restrict__ frags, const uint32_t *__restrict__ region, fc_params P,
                   ya_frag *__restrict__ work, fc_node *__restrict__ nodes, uint8_t *__restrict__ used, ya_frag *__restrict__ tmp,
                   ya_frag *__restrict__ path, ya_clump_rec *__restrict__ clumps, uint32_t *__restrict__ count,
                   uint32_t *__restrict__ first_out, uint32_t *__restrict__ slot_strand)
{
    // one WARP per strand (form_clumps.h with FC_WARP_COOP: lane 0 runs the serial parts, the chain DP's inner loop is strided
    // over the lanes).
    // slot_strand[i] (zeroed by the caller): strand + 1 if slot i of the clump array holds a clump -- lets the kernels behind
    // this one run one thread per CLUMP instead of one per strand
    const int s = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = (int)(threadIdx.x & 31u);
    if (s >= n_seg) return;
    const ya_strand_frags sf = strands[s];
    const uint32_t n = sf.n_frags;
    const uint32_t first = n ? sf.first : 0;
    if (lane == 0) first_out[s] = first;
    if (n == 0) { if (lane == 0) count[s] = 0; return; }
    if (n > kMaxStrandFrags) { if (lane == 0) count[s] = 0xFFFFFFFFu; return; }
    const int readLen = (int)(read_off[(s >> 1) + 1] - read_off[s >> 1]);
    ya_frag *w = work + first;
    for (uint32_t k = lane; k < n; k += 32) w[k] = frags[first + k];     // the graph edits fragments in place: work on a copy
    __syncwarp();
    const int nc = fc_form_clumps(&P, w, region + first, (int)n, readLen, nodes + first, used + 2 * (size_t)first, tmp + first,
                                  path + first, clumps + first);
    if (lane == 0) {
        for (int k = 0; k < nc; k++) { clumps[first + k].first += first; slot_strand[first + k] = (uint32_t)s + 1u; }   // path indices absolute
        count[s] = (uint32_t)nc;
    }
}

// device_only (ya_align_batch): records stay on the device, nothing is copied back and the call does not wait.
int ya_form_clumps_impl(ya_ctx *c, ya_clump_batch *out, bool device_only)
{
    if (!c || !out || (!device_only && (!out->clump_first || !out->clump_count))) return YA_E_ARG;
    YA_CUDA(c, cudaSetDevice(c->device));
    out->n_clumps = 0; out->n_path = 0;
    c->fc_valid = false;
    const int n_seg = 2 * c->n_reads;
    if (n_seg == 0) return YA_OK;
    if (c->seed_chunks != 1) return ya_fail(c, YA_E_STATE, "ya_form_clumps: needs the survivors of a single-chunk ya_seed_frags call on the device");
    const size_t nk = c->seed_nkeep;
    if (!device_only && (nk > out->cap || (nk && (!out->clumps || !out->path)))) return ya_fail(c, YA_E_CAPACITY, "clump output buffers too small");
    cudaStream_t st = c->stream;
    AllocScope allocScope(st);
    YA_CUDA(c, c->d_fc_count.reserve((size_t)n_seg * 8 + 64));     // counts, firsts and (tail) the job counter of ya_prepare_clumps
    uint32_t *d_count = c->d_fc_count.as<uint32_t>(), *d_first = d_count + n_seg;
    YA_CUDA(c, c->d_fc_work.reserve(nk * sizeof(ya_frag) + 64));
    YA_CUDA(c, c->d_fc_tmp.reserve(nk * sizeof(ya_frag) + 64));
    YA_CUDA(c, c->d_fc_path.reserve(nk * sizeof(ya_frag) + 64));
    YA_CUDA(c, c->d_fc_nodes.reserve(nk * sizeof(fc_node) + 64));
    YA_CUDA(c, c->d_fc_used.reserve(nk * 2 + 64));
    YA_CUDA(c, c->d_fc_clumps.reserve(nk * sizeof(ya_clump_rec) + 64));
    YA_CUDA(c, c->d_fc_slot.reserve(nk * 4 + 64));
    YA_CUDA(c, cudaMemsetAsync(c->d_fc_slot.p, 0, nk * 4 + 4, st));
    fc_params P;
    P.wordLen = c->P.wordLen; P.maxGap = c->P.maxGap; P.maxDesert = out->maxDesert; P.minMatch = c->P.minMatch;
    P.minNonOverlap = out->minNonOverlap; P.bandWidth = c->P.bandWidth; P.GOCost = c->P.GOCost; P.GECost = c->P.GECost; P.MScore = c->P.MScore;
    YA_CUDA(c, cudaEventRecord(c->ev[0], st));
    form_clumps_kernel<<<(n_seg + 3) / 4, 128, 0, st>>>(c->d_strand_out.as<ya_strand_frags>(), n_seg, c->d_read_off.as<uint64_t>(),
        c->d_frags_out.as<ya_frag>(), c->d_region_out.as<uint32_t>(), P, c->d_fc_work.as<ya_frag>(), c->d_fc_nodes.as<fc_node>(),
        c->d_fc_used.as<uint8_t>(), c->d_fc_tmp.as<ya_frag>(), c->d_fc_path.as<ya_frag>(), c->d_fc_clumps.as<ya_clump_rec>(), d_count, d_first,
        c->d_fc_slot.as<uint32_t>());
    c->ctr.launches++;
    YA_CUDA(c, cudaEventRecord(c->ev[1], st));
    if (device_only) { YA_CUDA(c, cudaGetLastError()); c->fc_valid = true; return YA_OK; }
    YA_CUDA(c, cudaMemcpyAsync(out->clump_count, d_count, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, cudaMemcpyAsync(out->clump_first, d_first, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    if (nk) {
        YA_CUDA(c, cudaMemcpyAsync(out->clumps, c->d_fc_clumps.p, nk * sizeof(ya_clump_rec), cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, cudaMemcpyAsync(out->path, c->d_fc_path.p, nk * sizeof(ya_frag), cudaMemcpyDeviceToHost, st));
    }
    YA_CUDA(c, ya_stream_wait(st));
    YA_CUDA(c, cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->ctr.ms_seed += ms;
    size_t nc = 0, np = 0;
    for (int s = 0; s < n_seg; s++) {
        const uint32_t k = out->clump_count[s];
        if (k == 0xFFFFFFFFu) continue;
        nc += k;
        for (uint32_t q = 0; q < k; q++) np += out->clumps[out->clump_first[s] + q].n;
    }
    out->n_clumps = nc; out->n_path = np;
    c->fc_valid = true;
    return YA_OK;
}

// ------------------------------------------------------------------------------------------------------------
// First phase of alignClump for those clumps (prepare_clumps.h), one thread per clump.
__global__ void __launch_bounds__(128)
prepare_clumps_kernel(uint32_t n_slots, const uint32_t *__restrict__ slot_strand, const uint64_t *__restrict__ read_off,
                      const ya_clump_rec *__restrict__ clumps, const ya_frag *__restrict__ path_in,
                      const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd, const uint8_t *__restrict__ rev, pc_params P,
                      ya_frag *__restrict__ path, ya_gap_rec *__restrict__ gaps, ya_prep_rec *__restrict__ prep,
                      ya_dp_job *__restrict__ jobs, uint32_t *__restrict__ n_jobs, uint32_t jobs_cap)
{
    // one thread per clump (slot_strand: form_clumps_kernel)
    // n_jobs[0]: job index allocator; n_jobs[1]: extension jobs among them (sizes the extension launch of the round)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const uint32_t sp1 = slot_strand[i];
    if (sp1 == 0) return;
    const uint32_t s = sp1 - 1, r = s >> 1;
    const int strand = (int)(s & 1u);
    const uint64_t base = read_off[r];
    const int readLen = (int)(read_off[r + 1] - base);
    const uint8_t *q = (strand ? rev : fwd) + base;
    const ya_clump_rec rec = clumps[i];
    for (uint32_t f = 0; f < rec.n; f++) path[rec.first + f] = path_in[rec.first + f];
    ya_prep_rec pr;
    pr.gap_first = rec.first;
    pc_prepare_clump(&P, bases, q, readLen, r, strand, path + rec.first, (int)rec.n, gaps + rec.first, jobs, n_jobs, jobs_cap, &pr);
    prep[i] = pr;
    const uint32_t ne = (pr.jobB != 0xFFFFFFFFu) + (pr.jobF != 0xFFFFFFFFu);
    if (ne) atomicAdd(n_jobs + 1, ne);
}

// device_only (ya_align_batch): only the two job counts come back (out->n_jobs, *n_ext); records and jobs stay on the device.
int ya_prepare_clumps_impl(ya_ctx *c, ya_prep_batch *out, bool device_only, size_t *n_ext)
{
    if (!c || !out) return YA_E_ARG;
    if (n_ext) *n_ext = 0;
    YA_CUDA(c, cudaSetDevice(c->device));
    out->n_jobs = 0;
    const int n_seg = 2 * c->n_reads;
    if (n_seg == 0) return YA_OK;
    if (!c->fc_valid) return ya_fail(c, YA_E_STATE, "ya_prepare_clumps: call ya_form_clumps first");
    const size_t nk = c->seed_nkeep;
    if (nk == 0) return YA_OK;
    if (!device_only && (nk > out->cap || !out->prep || !out->gaps || !out->path || !out->jobs)) return ya_fail(c, YA_E_CAPACITY, "prepare output buffers too small");
    cudaStream_t st = c->stream;
    AllocScope allocScope(st);
    const size_t jobsCap = device_only ? 3 * nk + 16 : std::min<size_t>(out->jobs_cap, 3 * nk + 16);
    YA_CUDA(c, c->d_pc_path.reserve(nk * sizeof(ya_frag) + 64));
    YA_CUDA(c, c->d_pc_gaps.reserve(nk * sizeof(ya_gap_rec) + 64));
    YA_CUDA(c, c->d_pc_prep.reserve(nk * sizeof(ya_prep_rec) + 64));
    YA_CUDA(c, c->d_pc_jobs.reserve(jobsCap * sizeof(ya_dp_job) + 64));
    YA_CUDA(c, c->h_stage3.reserve(64));
    uint32_t *d_count = c->d_fc_count.as<uint32_t>(), *d_first = d_count + n_seg;
    uint32_t *d_njobs = d_first + n_seg;                               // (d_fc_count has a spare tail)
    YA_CUDA(c, cudaMemsetAsync(d_njobs, 0, 8, st));
    pc_params P;
    P.bandWidth = c->P.bandWidth; P.GOCost = c->P.GOCost; P.GECost = c->P.GECost; P.RCost = c->P.RCost; P.MScore = c->P.MScore;
    P.minExtLength = c->P.minExtLength; P.maxROff = c->maxROff;
    YA_CUDA(c, cudaEventRecord(c->ev[0], st));
    prepare_clumps_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, st>>>((uint32_t)nk, c->d_fc_slot.as<uint32_t>(), c->d_read_off.as<uint64_t>(),
        c->d_fc_clumps.as<ya_clump_rec>(), c->d_fc_path.as<ya_frag>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), P,
        c->d_pc_path.as<ya_frag>(), c->d_pc_gaps.as<ya_gap_rec>(), c->d_pc_prep.as<ya_prep_rec>(), c->d_pc_jobs.as<ya_dp_job>(), d_njobs,
        (uint32_t)jobsCap);
    c->ctr.launches++;
    YA_CUDA(c, cudaEventRecord(c->ev[1], st));
    uint32_t *h_n = c->h_stage3.as<uint32_t>();
    YA_CUDA(c, cudaMemcpyAsync(h_n, d_njobs, 8, cudaMemcpyDeviceToHost, st));
    if (device_only) {
        YA_CUDA(c, ya_stream_wait(st));
        YA_CUDA(c, cudaGetLastError());
        float msd = 0; cudaEventElapsedTime(&msd, c->ev[0], c->ev[1]);
        c->ctr.ms_seed += msd;
        if (h_n[0] > jobsCap) return ya_fail(c, YA_E_STATE, "prepare: more DP jobs than the job buffer holds");
        out->n_jobs = h_n[0];
        if (n_ext) *n_ext = h_n[1];
        return YA_OK;
    }
    YA_CUDA(c, cudaMemcpyAsync(out->prep, c->d_pc_prep.p, nk * sizeof(ya_prep_rec), cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, cudaMemcpyAsync(out->gaps, c->d_pc_gaps.p, nk * sizeof(ya_gap_rec), cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, cudaMemcpyAsync(out->path, c->d_pc_path.p, nk * sizeof(ya_frag), cudaMemcpyDeviceToHost, st));
    // jobs: speculatively 8 per read, the rest after the count is known
    const size_t guess = std::min<size_t>(jobsCap, (size_t)8 * (size_t)c->n_reads + 64);
    YA_CUDA(c, cudaMemcpyAsync(out->jobs, c->d_pc_jobs.p, guess * sizeof(ya_dp_job), cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, ya_stream_wait(st));
    YA_CUDA(c, cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->ctr.ms_seed += ms;
    const size_t nj = *h_n;
    if (nj > jobsCap) return ya_fail(c, YA_E_CAPACITY, "prepare: more DP jobs than the job buffer holds");
    if (nj > guess) {
        YA_CUDA(c, cudaMemcpyAsync(out->jobs + guess, c->d_pc_jobs.as<ya_dp_job>() + guess, (nj - guess) * sizeof(ya_dp_job), cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, ya_stream_wait(st));
    }
    out->n_jobs = nj;
    return YA_OK;
}

extern "C" int ya_form_clumps(ya_ctx *c, ya_clump_batch *out) { return ya_form_clumps_impl(c, out, false); }
extern "C" int ya_prepare_clumps(ya_ctx *c, ya_prep_batch *out) { return ya_prepare_clumps_impl(c, out, false, nullptr); }
