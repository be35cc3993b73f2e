// sw.cu -- stage 3 of the alignment hot path: yaha's affine-gap DP family on the device.
//
// Replaces findAffineGapScore<banded,extension,reverse,XCutoff> (SW.cpp:798-1208) behind the four
// wrappers findAGSAlignment / findAGSAlignmentBanded / findAGSForwardExtension /
// findAGSBackwardExtension (SW.cpp:462-547), including decompressRef (SW.cpp:444-456).
//
// Kernels
//   dp_wave_kernel<G,C,EXT>  banded DP (X-drop extension or both-ends-anchored), one group of G
//                            lanes per job, C band columns per lane held in registers, rows flow
//                            through the lanes as a wavefront (lane l works on row m-l at macro
//                            step m); neighbour state moves with warp shuffles.  INT32-issue bound.
//   dp_thread_kernel         generic one-thread-per-job version of the same recurrences, used for
//                            the tiny full-matrix jobs and as the fall-back for band widths the
//                            wavefront kernel is not instantiated for.
//   traceback_kernel         one thread per job walks the back-pointer cells (SW.cpp:1138-1195)
//   finalize / compact       result assembly and op-list compaction
//
// All arithmetic is int32 exactly as in the reference; results are bit-identical by construction
// and checked against the oracle in tests/.
#include "common.cuh"
#include <algorithm>
#include <climits>

// back-pointer cell: op in bits 15..14 (0 M, 1 R, 2 D, 3 I), run length in bits 13..0
#define BP_M 0u
#define BP_R 1u
#define BP_D 2u
#define BP_I 3u

struct DpConst {
    int GOC, GEC, RC, MS, X, maxIntron, maxGap, BW;
};

__device__ __forceinline__ int nib(const uint8_t *__restrict__ bases, uint32_t off)
{
    uint32_t b = bases[off >> 1];
    return (off & 1) ? (b & 15u) : (b >> 4);          // Math.c:180-188
}

// One DP cell (SW.cpp:1017-1063).  (Vl,El,Dl) is the left neighbour's state on entry and this
// cell's state on exit; (Vu,Fu,Iu) the insert predecessor; Vd the diagonal predecessor.
template <bool EXT>
__device__ __forceinline__ void dp_cell(const DpConst &K, bool match, int Vd, int Vu, int Fu, int Iu,
                                        int &Vl, int &El, int &Dl, int &Fo, int &Io, uint32_t &bp)
{
    int v = Vd + (match ? K.MS : -K.RC);
    uint32_t op = match ? BP_M : BP_R, len = 1;
    int CE = El - K.GEC, NE = Vl - (K.GOC + K.GEC);
    bool contE = (CE >= NE) && (Dl + 1 <= K.maxIntron);
    int E = contE ? CE : NE;
    int D = contE ? Dl + 1 : 1;
    bool takeE = EXT ? (E >= v) : (E > v);
    if (takeE) { v = E; op = BP_D; len = (uint32_t)D; }
    int CF = Fu - K.GEC, NF = Vu - (K.GOC + K.GEC);
    bool contF = (CF >= NF) && (Iu + 1 <= K.maxGap);
    int F = contF ? CF : NF;
    int I = contF ? Iu + 1 : 1;
    bool takeF = EXT ? (F >= v) : (F > v);
    if (takeF) { v = F; op = BP_I; len = (uint32_t)I; }
    Vl = v; El = E; Dl = D; Fo = F; Io = I;
    bp = (op << 14) | len;
}

// ------------------------------------------------------------------------------------------
// Wavefront kernel.
// ------------------------------------------------------------------------------------------
template <int G, int C, bool EXT>
__global__ void __launch_bounds__(128)
dp_wave_kernel(const DevJob *__restrict__ jobs, const uint32_t *__restrict__ job_ids, int n_jobs,
               DevJobOut *__restrict__ outs, uint16_t *__restrict__ tb,
               const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd,
               const uint8_t *__restrict__ rev, DpConst K)
{
    const int gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const int gidx = gthread / G;
    const int l = threadIdx.x % G;                       // lane within the group
    const bool have = gidx < n_jobs;
    DevJob J;
    if (have) J = jobs[job_ids[gidx]];
    const int rows = have ? J.qLen : 0;
    const int lb = have ? J.lb : 0;
    const int W = have ? (J.lb + J.rb + 1) : 1;
    const int rLen = have ? J.rLen : 0;
    const bool bwd = have && J.kind == YA_DP_EXT_BWD;
    const uint8_t *codes = (have && J.strand) ? rev : fwd;
    const uint32_t qIdx = have ? J.qIdx : 0;
    const uint32_t rOff = have ? J.rOff : 0;
    uint16_t *mytb = tb + (have ? J.tb_off : 0);

    int Vp[C], Fp[C], Ip[C], rc[C];
    const int col0 = l * C;
#pragma unroll
    for (int k = 0; k < C; k++) {
        int col = col0 + k;
        Fp[k] = YA_WORST; Ip[k] = 0;
        if (col == lb) { Vp[k] = 0; Fp[k] = 0; }                                   // origin, SW.cpp:914-916
        else if (col > lb && col < W) Vp[k] = -(K.GOC + (col - lb) * K.GEC);       // leading deletes, :900-910
        else Vp[k] = YA_WORST;                                                     // unread / sentinel, :894
    }
    // reference characters for the row this lane will look at next (row 1 - l at macro step 1)
    int ridx0 = (1 - l) - lb - 1 + col0;
#pragma unroll
    for (int k = 0; k < C; k++) {
        int ri = ridx0 + k;
        rc[k] = (ri >= 0 && ri < rLen) ? nib(bases, bwd ? rOff - (uint32_t)ri : rOff + (uint32_t)ri) : 0xFF;
    }
    int qi = 1 - l;                                       // row handled at the coming macro step
    int qc = (qi >= 1 && qi <= rows) ? codes[bwd ? qIdx - (uint32_t)(qi - 1) : qIdx + (uint32_t)(qi - 1)] : 0xFE;

    // state published to the right neighbour: last column of the row finished in the previous step
    int pubV = YA_WORST, pubE = YA_WORST, pubD = 0, pubMax = YA_WORST, pubArg = 0;
    // extension bookkeeping, authoritative on the last lane of the group
    int maxScore = YA_WORST, maxi = 0, maxj = 0;
    int stopRow = rows;                                   // last row that counts
    const unsigned full = 0xffffffffu;

    for (int m = 1;; m++) {
        const bool group_done = m > stopRow + G - 1;
        if (__all_sync(full, group_done)) break;
        const int i = m - l;
        const bool rowActive = (i >= 1) && (i <= stopRow);

        // 1. left neighbour's state for this row
        int Vl = __shfl_up_sync(full, pubV, 1, G);
        int El = __shfl_up_sync(full, pubE, 1, G);
        int Dl = __shfl_up_sync(full, pubD, 1, G);
        int rowMax = YA_WORST, rowArg = 0;
        if (EXT) {
            rowMax = __shfl_up_sync(full, pubMax, 1, G);
            rowArg = __shfl_up_sync(full, pubArg, 1, G);
        }
        if (l == 0) { Vl = YA_WORST; El = YA_WORST; Dl = 0; rowMax = YA_WORST; rowArg = 0; }

        int startCol = lb + 1 - i;                        // SW.cpp:975-981
        const int bcol = startCol - 1;                    // boundary cell (leading insert), if >= 0
        if (startCol < 0) startCol = 0;
        int endCol = lb + rLen - i; if (endCol > W - 1) endCol = W - 1;    // SW.cpp:983
        const int bV = -(K.GOC + i * K.GEC);

        // prefetch next row's query / reference characters
        const int qn = i + 1;
        int qc_next = (qn >= 1 && qn <= rows) ? codes[bwd ? qIdx - (uint32_t)(qn - 1) : qIdx + (uint32_t)(qn - 1)] : 0xFE;
        const int rn = ridx0 + C;                         // index of the char entering the window
        int rc_next = (rn >= 0 && rn < rLen) ? nib(bases, bwd ? rOff - (uint32_t)rn : rOff + (uint32_t)rn) : 0xFF;

        uint16_t *tbrow = mytb + (size_t)m * (G * C) + col0;

        int V0 = 0, F0 = 0, I0 = 0;                       // right neighbour's first column (row i-1)
#pragma unroll
        for (int k = 0; k < C; k++) {
            const int col = col0 + k;
            if (k == C - 1) {
                // 3. insert predecessor of the last column lives in the right neighbour, which has
                //    just finished its first column of row i-1 in this same macro step
                V0 = __shfl_down_sync(full, Vp[0], 1, G);
                F0 = __shfl_down_sync(full, Fp[0], 1, G);
                I0 = __shfl_down_sync(full, Ip[0], 1, G);
                if (l == G - 1) { V0 = YA_WORST; F0 = YA_WORST; I0 = 0; }          // sentinel column
            }
            const int Vu = (k == C - 1) ? V0 : Vp[(k + 1) % C];
            const int Fu = (k == C - 1) ? F0 : Fp[(k + 1) % C];
            const int Iu = (k == C - 1) ? I0 : Ip[(k + 1) % C];
            if (rowActive) {
                if (col >= startCol && col <= endCol) {
                    int Fo, Io; uint32_t bp;
                    dp_cell<EXT>(K, qc == rc[k], Vp[k], Vu, Fu, Iu, Vl, El, Dl, Fo, Io, bp);
                    Vp[k] = Vl; Fp[k] = Fo; Ip[k] = Io;
                    tbrow[k] = (uint16_t)bp;
                    if (EXT && Vl > rowMax) { rowMax = Vl; rowArg = col; }          // SW.cpp:1069
                } else if (col == bcol) {
                    Vp[k] = bV; Vl = bV; El = YA_WORST; Dl = 0;                     // SW.cpp:981, 965-966
                }
            }
        }
        pubV = Vl; pubE = El; pubD = Dl; pubMax = rowMax; pubArg = rowArg;

        // 4. row-major argmax and X-drop, decided by the last lane (SW.cpp:1073-1078, 1091)
        if (EXT) {
            int newStop = stopRow;
            if (l == G - 1 && rowActive) {
                if (rowMax > maxScore) { maxScore = rowMax; maxi = i; maxj = rowArg; }
                if (rowMax < maxScore - K.X) newStop = i;
            }
            stopRow = __shfl_sync(full, newStop, G - 1, G);
        }
        // slide the character windows
#pragma unroll
        for (int k = 0; k < C - 1; k++) rc[k] = rc[k + 1];
        rc[C - 1] = rc_next;
        ridx0++;
        qc = qc_next;
    }

    if (!have) return;
    if (EXT) {
        if (l == G - 1) {
            DevJobOut o;
            o.score = maxScore; o.maxi = maxi; o.maxj = maxj; o.n_ops = 0;
            o.cells_lo = (uint32_t)stopRow; o.cells_hi = 0;      // rows executed; cells derived later
            outs[job_ids[gidx]] = o;
        }
    } else {
        const int rbcol = J.rb;
        if (rbcol / C == l) {
            DevJobOut o;
            int v = 0;
#pragma unroll
            for (int k = 0; k < C; k++) if (k == rbcol % C) v = Vp[k];
            o.score = v; o.maxi = rows; o.maxj = rbcol; o.n_ops = 0;
            o.cells_lo = (uint32_t)rows; o.cells_hi = 0;
            outs[job_ids[gidx]] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Packed X-drop extension kernel (the hot kernel): same wavefront as dp_wave_kernel, but every
// score lives in a register scaled by 256 with the low byte carrying what the reference keeps in
// side arrays:
//     E-chain word = E*256 + 0x40 + (D-1)      F-chain word = F*256 + 0x80 + (I-1)
// A continued gap is "word + 1 - GEC*256", a new gap "V*256 + tag - (GOC+GEC)*256", so ONE
// VIADDMNMX yields the better of the two with the reference's tie rule (continue on equality,
// SW.cpp:1032,1050) and the run length D / I for free.  The cell value is ONE VIMNMX3 over
// {diag, E word, F word}: equal scores resolve by tag, i.e. insert over delete over match/replace,
// exactly the extension tie rule (SW.cpp:1036,1054), and the low byte of the winner IS the
// back-pointer (op tag + run length - 1).  Row maximum and its first column ride in one key
// V*256 + (255 - col).  Valid while run lengths fit 6 bits and the gap-length caps cannot bind
// (W <= 64, W <= maxGap, W <= maxIntron); other parameter sets use dp_wave_kernel.
// Back-pointer bytes are accumulated 4 macro steps per 32-bit word and stored 16 B at a time.
// ------------------------------------------------------------------------------------------
// ---- reference-window staging (north_star: "reference windows staged into shared memory via TMA"; replaces decompressRef,
// SW.cpp:444-456).  A job's window is rLen = rows + 2*BW bases = a few hundred bytes of the packed genome.  With STAGE the
// first lane of every warp issues one 1-D bulk copy (cp.async.bulk, the TMA engine without a tensor map) per job of the warp
// into the warp's shared-memory slots and all lanes wait on the warp's mbarrier; the DP loop then reads its one reference
// nibble per macro step from shared memory instead of the L1/L2 path.  Windows above kStageBytes stay on the global path.
#define kStageBytes 512
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

#define PK_WORST (-(1 << 29))
#define PK_TAGE  0x40
#define PK_TAGF  0x80

template <int G, int C, int W, bool STAGE = false>
__global__ void __launch_bounds__(128)
dp_ext_packed_kernel(const DevJob *__restrict__ jobs, const uint32_t *__restrict__ job_ids, int n_jobs,
                     DevJobOut *__restrict__ outs, uint32_t *__restrict__ tbw,
                     const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd,
                     const uint8_t *__restrict__ rev, DpConst K)
{
    constexpr int GPW = 32 / G;                  // groups (jobs) per warp
    constexpr int LB = (W - 1) / 2;              // left band = 2*BW
    constexpr int CP = (C + 3) & ~3;             // words reserved per lane and 4-step block (16 B aligned)
    constexpr int KB = W - (G - 1) * C;          // first blocked column index in the last lane
    static_assert(G * C >= W && (G - 1) * C < W, "last lane must hold the band's right edge");
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int g = lane / G, l = lane % G;
    const bool laneUsed = g < GPW;
    const int gidx = warpGlobal * GPW + g;
    const bool have = laneUsed && gidx < n_jobs;
    const int leftLane = (l > 0) ? lane - 1 : lane;
    const int rightLane = (l < G - 1 && lane < 31) ? lane + 1 : lane;
    const int lastLane = g * G + (G - 1);
    uint32_t jid = 0;
    DevJob J;
    if (have) { jid = job_ids[gidx]; J = jobs[jid]; }
    const int rows = have ? J.qLen : 0;
    const int rLen = have ? J.rLen : 0;
    const bool bwd = have && J.kind == YA_DP_EXT_BWD;
    const uint8_t *codes = (have && J.strand) ? rev : fwd;
    const uint32_t qIdx = have ? J.qIdx : 0;
    const uint32_t rOff = have ? J.rOff : 0;
    uint32_t *mytb = tbw + (have ? J.tb_off : 0);

    // ---- STAGE: the warp's reference windows into shared memory by bulk copies
    __shared__ __align__(16) uint8_t s_win[STAGE ? 4 : 1][STAGE ? GPW : 1][STAGE ? kStageBytes : 16];
    __shared__ __align__(8) uint64_t s_bar[4];
    const uint8_t *win = nullptr;                // this job's staged window (byte `winBase` of the genome at win[0]), or null
    uint32_t winBase = 0;
    if (STAGE) {
        const int wib = threadIdx.x >> 5;
        // byte range of the window, widened to 16-byte granules (the genome array is padded: ctx.cu)
        const uint32_t loB = have ? (bwd ? rOff - (uint32_t)(rLen - 1) : rOff) : 0u;
        const uint32_t a0 = (loB >> 1) & ~15u;
        const uint32_t nby = have ? ((((loB + (uint32_t)rLen - 1u) >> 1) - a0 + 1u + 15u) & ~15u) : 0u;
        const bool fits = have && rLen > 0 && nby <= (uint32_t)kStageBytes;
        if (lane == 0) mbar_init(&s_bar[wib], 1);
        __syncwarp();
        uint32_t total = 0;
#pragma unroll
        for (int q = 0; q < GPW; q++) {
            const uint32_t nb = __shfl_sync(full, fits ? nby : 0u, q * G);
            total += nb;
        }
        if (lane == 0 && total) mbar_expect_tx(&s_bar[wib], total);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < GPW; q++) {
            const uint32_t nb = __shfl_sync(full, fits ? nby : 0u, q * G);
            const uint32_t aq = __shfl_sync(full, a0, q * G);
            if (lane == 0 && nb) bulk_g2s(&s_win[wib][q][0], bases + aq, nb, &s_bar[wib]);
        }
        bool arrived = total == 0;
        for (int spin = 0; !arrived && spin < (1 << 20); spin++) arrived = mbar_try_wait(&s_bar[wib], 0);
        arrived = __all_sync(full, arrived);     // (a copy that never lands would only cost the staging: fall back to global loads)
        if (fits && arrived && laneUsed) { win = &s_win[wib][g][0]; winBase = a0; }
    }
    auto refc = [&](uint32_t off) -> int {
        if (STAGE && win) { const uint32_t b = win[(off >> 1) - winBase]; return (off & 1) ? (int)(b & 15u) : (int)(b >> 4); }
        return nib(bases, off);
    };

    const int MS8 = K.MS << 8, RC8 = K.RC << 8;
    const int CONT = 1 - (K.GEC << 8);
    const int NEWE = PK_TAGE - ((K.GOC + K.GEC) << 8), NEWF = PK_TAGF - ((K.GOC + K.GEC) << 8);
    const int E_NONE = PK_WORST + PK_TAGE - 1;    // "no gap yet" (D = 0): continuing it gives D = 1
    const int F_NONE = PK_WORST + PK_TAGF - 1;

    int V[C], F[C], rc[C];
    uint32_t acc[C];
    const int col0 = l * C;
#pragma unroll
    for (int k = 0; k < C; k++) {
        const int col = col0 + k;
        F[k] = F_NONE; acc[k] = 0;
        if (col == LB) V[k] = 0;                                                  // origin (SW.cpp:914-916)
        else if (col > LB && col < W) V[k] = -((K.GOC + (col - LB) * K.GEC) << 8);  // leading deletes (:900-910)
        else V[k] = PK_WORST;                                                     // unread / sentinel (:894)
    }
    int ridx0 = (1 - l) - LB - 1 + col0;
#pragma unroll
    for (int k = 0; k < C; k++) {
        const int ri = ridx0 + k;
        rc[k] = (ri >= 0 && ri < rLen) ? refc(bwd ? rOff - (uint32_t)ri : rOff + (uint32_t)ri) : 0xFF;
    }
    int qi = 1 - l;
    int qc = (qi >= 1 && qi <= rows) ? codes[bwd ? qIdx - (uint32_t)(qi - 1) : qIdx + (uint32_t)(qi - 1)] : 0xFE;

    int pubV = PK_WORST, pubE = E_NONE, pubRm = INT_MIN;
    int maxScore = YA_WORST, maxi = 0, maxj = 0;
    int stopRow = rows;
    const int kbase = 255 - col0;

    int m = 1;
    for (;; m++) {
        const bool group_done = m > stopRow + G - 1;
        if (__all_sync(full, group_done)) break;
        const int i = m - l;
        const bool rowActive = (i >= 1) && (i <= stopRow);

        int Vl = __shfl_sync(full, pubV, leftLane);
        int El = __shfl_sync(full, pubE, leftLane);
        int rm = __shfl_sync(full, pubRm, leftLane);
        if (l == 0) { Vl = PK_WORST; El = E_NONE; rm = INT_MIN; }

        const int qn = i + 1;
        const int qc_next = (qn >= 1 && qn <= rows) ? codes[bwd ? qIdx - (uint32_t)(qn - 1) : qIdx + (uint32_t)(qn - 1)] : 0xFE;
        const int rn = ridx0 + C;
        const int rc_next = (rn >= 0 && rn < rLen) ? refc(bwd ? rOff - (uint32_t)rn : rOff + (uint32_t)rn) : 0xFF;

        const bool general = m < LB + G;          // some lane of the group may still be inside the leading triangle
        int V0r = PK_WORST, F0r = F_NONE;
        if (general) {
            // ---- leading triangle: per-cell activity tests (rows 1..LB have a moving left edge)
            const int startCol = (LB + 1 - i) > 0 ? (LB + 1 - i) : 0;
            const int bcol = LB - i;
            const int bV = -((K.GOC + i * K.GEC) << 8);
#pragma unroll
            for (int k = 0; k < C; k++) {
                const int col = col0 + k;
                if (k == 1) {
                    V0r = __shfl_sync(full, V[0], rightLane);
                    F0r = __shfl_sync(full, F[0], rightLane);
                    if (l == G - 1) { V0r = PK_WORST; F0r = F_NONE; }
                }
                const bool blocked = (k >= KB) && (l == G - 1);
                bool doCell = rowActive && !blocked;
                if (doCell && col < startCol) {
                    doCell = false;
                    if (col == bcol) { V[k] = bV; Vl = bV; El = E_NONE; }         // leading-insert boundary (SW.cpp:981)
                }
                if (doCell) {
                    const int Vu = (k == C - 1) ? V0r : V[(k + 1) % C];
                    const int Fu = (k == C - 1) ? F0r : F[(k + 1) % C];
                    const int diag = V[k] + ((qc == rc[k]) ? MS8 : -RC8);
                    const int Et = __viaddmax_s32(El, CONT, Vl + NEWE);
                    const int Ft = __viaddmax_s32(Fu, CONT, Vu + NEWF);
                    const int r = __vimax3_s32(diag, Et, Ft);
                    const int Vn = r & ~255;
                    acc[k] = acc[k] * 256u + (uint32_t)(r - Vn);
                    rm = max(rm, Vn + kbase - k);
                    V[k] = Vn; F[k] = Ft; Vl = Vn; El = Et;
                } else {
                    acc[k] = acc[k] * 256u;
                }
            }
        } else {
            // ---- steady state: straight-line code, no per-cell tests.  Lanes whose row is outside
            // 1..stopRow compute on dead registers (nothing reads them any more); only the
            // sentinel column right of the band (cell KB of the last lane) must be kept intact,
            // and blocked cells must not enter the row maximum.
            const bool lastL = (l == G - 1);
#pragma unroll
            for (int k = 0; k < C; k++) {
                if (k == 1) {
                    V0r = __shfl_sync(full, V[0], rightLane);
                    F0r = __shfl_sync(full, F[0], rightLane);
                    if (lastL) { V0r = PK_WORST; F0r = F_NONE; }
                }
                const int Vu = (k == C - 1) ? V0r : V[(k + 1) % C];
                const int Fu = (k == C - 1) ? F0r : F[(k + 1) % C];
                const int diag = V[k] + ((qc == rc[k]) ? MS8 : -RC8);
                const int Et = __viaddmax_s32(El, CONT, Vl + NEWE);
                const int Ft = __viaddmax_s32(Fu, CONT, Vu + NEWF);
                const int r = __vimax3_s32(diag, Et, Ft);
                const int Vn = r & ~255;
                acc[k] = acc[k] * 256u + (uint32_t)(r - Vn);
                if (k >= KB) {
                    rm = lastL ? rm : max(rm, Vn + kbase - k);
                    if (k == KB) { V[k] = lastL ? PK_WORST : Vn; F[k] = lastL ? F_NONE : Ft; }
                    else { V[k] = Vn; F[k] = Ft; }
                } else {
                    rm = max(rm, Vn + kbase - k);
                    V[k] = Vn; F[k] = Ft;
                }
                Vl = Vn; El = Et;
            }
        }
        pubV = Vl; pubE = El; pubRm = rm;

        if ((m & 3) == 0) {                       // one 16 B-aligned burst per lane every 4 macro steps
            uint32_t *dst = mytb + (size_t)((m >> 2) - 1) * (G * CP) + l * CP;
            if (have && m <= rows + G + 3) {
#pragma unroll
                for (int k = 0; k < CP; k += 4) {
                    uint4 w;
                    w.x = acc[k < C ? k : 0]; w.y = (k + 1 < C) ? acc[k + 1] : 0u;
                    w.z = (k + 2 < C) ? acc[k + 2] : 0u; w.w = (k + 3 < C) ? acc[k + 3] : 0u;
                    *reinterpret_cast<uint4 *>(dst + k) = w;
                }
            }
        }

        int newStop = stopRow;
        if (lane == lastLane && rowActive) {      // row-major argmax + X-drop (SW.cpp:1073-1078, 1091)
            const int rV = rm >> 8;
            if (rV > maxScore) { maxScore = rV; maxi = i; maxj = 255 - (rm & 255); }
            if (rV < maxScore - K.X) newStop = i;
        }
        stopRow = __shfl_sync(full, newStop, laneUsed ? lastLane : lane);
#pragma unroll
        for (int k = 0; k < C - 1; k++) rc[k] = rc[k + 1];
        rc[C - 1] = rc_next;
        ridx0++;
        qc = qc_next;
    }
    // flush the partial 4-step block (m is one past the last executed macro step)
    {
        const int done = m - 1;
        if ((done & 3) != 0 && have && done <= rows + G + 3) {
            const int sh = 8 * (4 - (done & 3));
            uint32_t *dst = mytb + (size_t)(done >> 2) * (G * CP) + l * CP;
#pragma unroll
            for (int k = 0; k < C; k++) dst[k] = acc[k] << sh;
        }
    }
    if (have && lane == lastLane) {
        DevJobOut o;
        o.score = maxScore; o.maxi = maxi; o.maxj = maxj; o.n_ops = 0;
        o.cells_lo = (uint32_t)stopRow; o.cells_hi = 0;
        outs[jid] = o;
    }
}

// ------------------------------------------------------------------------------------------
// Generic one-thread-per-job kernel (all four kinds).  Row state lives in global scratch.
// FULL jobs are expressed in band coordinates with lb = qLen, rb = rLen, which reproduces the
// full-matrix recurrences, boundaries and traceback of SW.cpp exactly.
// ------------------------------------------------------------------------------------------
__global__ void dp_thread_kernel(const DevJob *__restrict__ jobs, const uint32_t *__restrict__ job_ids, int n_jobs,
                                 DevJobOut *__restrict__ outs, uint16_t *__restrict__ tb, int *__restrict__ rowbuf,
                                 const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd,
                                 const uint8_t *__restrict__ rev, DpConst K)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    const DevJob J = jobs[job_ids[t]];
    const bool ext = J.kind >= YA_DP_EXT_FWD, bwd = J.kind == YA_DP_EXT_BWD;
    const int rows = J.qLen, lb = J.lb, W = J.lb + J.rb + 1, rLen = J.rLen;
    const uint8_t *codes = J.strand ? rev : fwd;
    int *Vp = rowbuf + (size_t)J.rows_off, *Fp = Vp + (W + 1), *Ip = Fp + (W + 1);
    uint16_t *mytb = tb + J.tb_off;
    // row 0
    for (int j = 0; j <= W; j++) { Vp[j] = YA_WORST; Fp[j] = YA_WORST; Ip[j] = 0; }
    Vp[lb] = 0; Fp[lb] = 0;
    for (int j = lb + 1, k = 1; j < W; j++, k++) Vp[j] = -(K.GOC + k * K.GEC);
    int maxScore = YA_WORST, maxi = 0, maxj = 0, lastV = 0, doneRows = rows;
    for (int i = 1; i <= rows; i++) {
        int startCol = lb + 1 - i, Vl, El = YA_WORST, Dl = 0;
        if (startCol <= 0) { startCol = 0; Vl = YA_WORST; }
        else Vl = -(K.GOC + i * K.GEC);
        int endCol = lb + rLen - i; if (endCol > W - 1) endCol = W - 1;
        const int qc = codes[bwd ? J.qIdx - (uint32_t)(i - 1) : J.qIdx + (uint32_t)(i - 1)];
        int rowMax = YA_WORST;
        // in-place update: column j reads old [j] (diag) and old [j+1] (insert) before writing [j]
        // the boundary cell becomes the next row's diagonal predecessor
        int saveBoundary = Vl;
        for (int j = startCol; j <= endCol; j++) {
            const int ri = i - lb - 1 + j;
            const int rcode = nib(bases, bwd ? J.rOff - (uint32_t)ri : J.rOff + (uint32_t)ri);
            int Vd = Vp[j], Vu = Vp[j + 1], Fu = Fp[j + 1], Iu = Ip[j + 1], Fo, Io; uint32_t bp;
            if (ext) dp_cell<true>(K, qc == rcode, Vd, Vu, Fu, Iu, Vl, El, Dl, Fo, Io, bp);
            else     dp_cell<false>(K, qc == rcode, Vd, Vu, Fu, Iu, Vl, El, Dl, Fo, Io, bp);
            Vp[j] = Vl; Fp[j] = Fo; Ip[j] = Io;
            mytb[(size_t)i * J.stride + j] = (uint16_t)bp;
            if (Vl > rowMax) rowMax = Vl;
            if (ext && Vl > maxScore) { maxScore = Vl; maxi = i; maxj = j; }
            lastV = Vl;
        }
        if (startCol > 0) Vp[startCol - 1] = saveBoundary;
        if (ext && rowMax < maxScore - K.X) { doneRows = i; break; }
    }
    DevJobOut o;
    o.score = ext ? maxScore : lastV;
    o.maxi = ext ? maxi : rows;
    o.maxj = ext ? maxj : J.rb;
    o.n_ops = 0; o.cells_lo = (uint32_t)doneRows; o.cells_hi = 0;
    outs[job_ids[t]] = o;
}

// ------------------------------------------------------------------------------------------
// Traceback: one thread per job (SW.cpp:1138-1195).  Emits runs in walking order (end -> start).
// Row 0 and the leading-insert boundary cells are not stored; they are known in closed form
// (SW.cpp:900-933).
// ------------------------------------------------------------------------------------------
// Cells the reference executed in rows 1..done of a band (SW.cpp:1007 loop bounds):
//   sum_i max(0, min(W-1, lb+rLen-i) - max(0, lb+1-i) + 1),  piecewise linear in i with kinks at
//   i = lb+1 and i = rLen-rb, zero beyond i = lb+rLen -- summed per linear piece.
__device__ __forceinline__ uint64_t band_cells(int done, int lb, int rb, int rLen)
{
    const int W = lb + rb + 1;
    int hi = min(done, lb + rLen);
    if (hi < 1) return 0;
    auto f = [&](int i) { return (long long)(min(W - 1, lb + rLen - i) - max(0, lb + 1 - i) + 1); };
    int k1 = min(lb + 1, rLen - rb), k2 = max(lb + 1, rLen - rb);
    uint64_t total = 0;
    int lo = 1;
    const int cuts[3] = {k1, k2, hi};
    for (int q = 0; q < 3; q++) {
        int b = min(cuts[q], hi);
        if (b >= lo) { total += (uint64_t)((f(lo) + f(b)) * (long long)(b - lo + 1) / 2); lo = b + 1; }
    }
    return total;
}

// One thread per job; thread t takes job ids[t] -- the kernel-class / length order the fill kernels
// were launched in, so the lanes of a warp walk paths of the same layout and similar length.
__global__ void traceback_kernel(const DevJob *__restrict__ jobs, const uint32_t *__restrict__ ids, int n_jobs,
                                 DevJobOut *__restrict__ outs,
                                 const uint16_t *__restrict__ tb, ya_op *__restrict__ ops_raw,
                                 const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd,
                                 const uint8_t *__restrict__ rev)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    t = (int)ids[t];
    const DevJob J = jobs[t];
    DevJobOut o = outs[t];
    const bool ext = J.kind >= YA_DP_EXT_FWD;
    const int lb = J.lb, W = J.lb + J.rb + 1;
    {
        const uint64_t cells = band_cells((int)o.cells_lo, lb, (int)J.rb, (int)J.rLen);
        o.cells_lo = (uint32_t)cells; o.cells_hi = (uint32_t)(cells >> 32);
    }
    uint32_t n = 0;
    if (!(ext && o.score <= 0)) {
        const uint16_t *mytb = tb + (J.layout == 2 ? 0 : J.tb_off);
        ya_op *out = ops_raw + J.ops_off;
        int y = o.maxi, x = o.maxj;
        int prev = -1; uint32_t run = 0;
        size_t cachedIdx = (size_t)-1, nextIdx = (size_t)-1; uint32_t cachedWord = 0, nextWord = 0;
        int guard = (int)J.qLen + (int)J.rLen + 8;     // a valid walk consumes a row or a column per step
        for (;;) {
            if (--guard < 0 || x < 0 || x > W || y < 0) { n = 0xFFFFFFF0u; prev = -1; break; }   // corrupt back-pointers
            uint32_t op, len;
            if (y == 0) {
                if (x == lb) break;                      // origin, 'U'
                op = BP_D; len = (uint32_t)(x - lb);
            } else if (x == lb - y) {
                op = BP_I; len = (uint32_t)y;            // leading insert boundary
            } else if (J.layout == 2) {
                // packed bytes, 4 macro steps per word (dp_ext_packed_kernel); M vs R is not stored
                const int C = J.colsPerLane, CP = (C + 3) & ~3;
                const int s = y + x / C;
                const uint32_t *w = reinterpret_cast<const uint32_t *>(tb) + J.tb_off;
                // a word holds 4 consecutive macro steps of one column: a run of match/replace steps
                // (same column, rows y, y-1, ...) is served by one load per 4 rows
                const size_t widx = (size_t)((s - 1) >> 2) * J.stride + (x / C) * CP + (x % C);
                if (widx != cachedIdx) {
                    // the walk is latency bound: the word one block up in the same column (where a
                    // match/replace run continues) is requested together with the one needed now
                    cachedWord = (widx == nextIdx) ? nextWord : w[widx];
                    cachedIdx = widx;
                    if (widx >= J.stride) { nextIdx = widx - J.stride; nextWord = w[nextIdx]; } else nextIdx = (size_t)-1;
                }
                const uint32_t word = cachedWord;
                const uint32_t b = (word >> (8 * (3 - ((s - 1) & 3)))) & 0xFFu;
                len = (b & 63u) + 1u;
                if ((b >> 6) == 1u) op = BP_D;
                else if ((b >> 6) == 2u) op = BP_I;
                else {
                    const bool bwd = J.kind == YA_DP_EXT_BWD;
                    const uint8_t *codes = J.strand ? rev : fwd;
                    const int qc = codes[bwd ? J.qIdx - (uint32_t)(y - 1) : J.qIdx + (uint32_t)(y - 1)];
                    const int ri = y - lb - 1 + x;
                    const int rcode = nib(bases, bwd ? J.rOff - (uint32_t)ri : J.rOff + (uint32_t)ri);
                    op = (qc == rcode) ? BP_M : BP_R;
                }
            } else {
                size_t row = J.layout ? (size_t)(y + x / J.colsPerLane) : (size_t)y;
                uint32_t c = mytb[row * J.stride + x];
                op = c >> 14; len = c & 0x3FFFu;
            }
            if (op == BP_D) x -= (int)len;
            else if (op == BP_I) { y -= (int)len; x += (int)len; }
            else { y -= 1; len = 1; }
            if ((int)op != prev) {
                if (prev >= 0) {
                    if (n < J.ops_cap) { out[n].length = (uint16_t)run; out[n].opcode = "MRDI"[prev]; out[n].pad = 0; }
                    n++;
                }
                prev = (int)op; run = len;
            } else run += len;
        }
        if (prev >= 0) {
            if (n < J.ops_cap) { out[n].length = (uint16_t)run; out[n].opcode = "MRDI"[prev]; out[n].pad = 0; }
            n++;
        }
    }
    o.n_ops = n;
    outs[t] = o;
}

// Warp-cooperative traceback: one warp per job.  The walk itself is sequential, but its two costs are not:
//   * back-pointer fetches -- lane k reads the cell k rows further up the current column, so a stretch of up to
//     32 match/replace steps (which stay in one band column) costs ONE round of loads instead of 32 dependent ones;
//   * match vs replace along that stretch (the packed layout does not store it) -- 32 query/reference code pairs
//     are compared at once and the mismatch mask is cut into runs with bit scans.
// Gap steps (one cell = one whole run of the gap, its length is in the cell) take one iteration each.
// Emits exactly the runs of traceback_kernel (same walking order, same merging of equal neighbours).
__global__ void __launch_bounds__(128)
traceback_warp_kernel(const DevJob *__restrict__ jobs, const uint32_t *__restrict__ ids, int n_jobs,
                      DevJobOut *__restrict__ outs, const uint16_t *__restrict__ tb, ya_op *__restrict__ ops_raw,
                      const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd, const uint8_t *__restrict__ rev,
                      int BW, ya_dp_result *__restrict__ res, uint32_t *__restrict__ ops_cnt,
                      unsigned long long *__restrict__ acct)
{
    // acct != nullptr (device rounds, ya_align_batch): every job's runs are left IN PLACE in genome order -- walking order is
    // end -> start, so forward and global jobs fill their slot from its end -- res[t].ops_off addresses ops_raw directly (no
    // scan / compaction pass), and the counters the host would add up are accumulated here:
    // acct[0] cells, acct[1] cells of packed-layout jobs, acct[2] error flags (1 corrupt back-pointers, 2 slot overflow)
    const bool inPlace = acct != nullptr;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = (int)(threadIdx.x & 31);
    if (warp >= n_jobs) return;
    const unsigned full = 0xffffffffu;
    const int t = (int)ids[warp];
    const DevJob J = jobs[t];
    DevJobOut o = outs[t];
    const bool ext = J.kind >= YA_DP_EXT_FWD;
    const int lb = J.lb, W = J.lb + J.rb + 1;
    {
        const uint64_t cells = band_cells((int)o.cells_lo, lb, (int)J.rb, (int)J.rLen);
        o.cells_lo = (uint32_t)cells; o.cells_hi = (uint32_t)(cells >> 32);
    }
    uint32_t n = 0;
    if (!(ext && o.score <= 0)) {
        ya_op *out = ops_raw + J.ops_off;
        const int layout = J.layout, C = J.colsPerLane, CP = (C + 3) & ~3;
        const uint16_t *mytb = tb + (layout == 2 ? 0 : J.tb_off);
        const uint32_t *w32 = reinterpret_cast<const uint32_t *>(tb) + J.tb_off;       // layout 2 only
        const bool bwd = J.kind == YA_DP_EXT_BWD;
        const uint8_t *codes = J.strand ? rev : fwd;
        int y = o.maxi, x = o.maxj;
        int prev = -1; uint32_t run = 0;
        int guard = (int)J.qLen + (int)J.rLen + 8;
        const bool fromEnd = inPlace && J.kind != YA_DP_EXT_BWD;            // (backward extensions walk in genome order, SW.cpp:1184-1185)
        auto put = [&](uint32_t k, int op, uint32_t len) {
            if (lane == 0 && k < J.ops_cap) {
                ya_op *o = fromEnd ? &out[J.ops_cap - 1 - k] : &out[k];
                o->length = (uint16_t)len; o->opcode = "MRDI"[op]; o->pad = 0;
            }
        };
        auto emit = [&](int op, uint32_t len) {
            if (op != prev) {
                if (prev >= 0) { put(n, prev, run); n++; }
                prev = op; run = len;
            } else run += len;
        };
        for (;;) {
            if (--guard < 0 || x < 0 || x > W || y < 0 || x < lb - y) { n = 0xFFFFFFF0u; prev = -1; break; }   // corrupt back-pointers
            if (y == 0) {
                if (x == lb) break;                                      // origin
                emit((int)BP_D, (uint32_t)(x - lb)); x = lb;             // leading deletes (row 0 is not stored)
                continue;
            }
            if (x == lb - y) { emit((int)BP_I, (uint32_t)y); x += y; y = 0; continue; }     // leading-insert boundary
            // lane k looks at the cell k rows up in this band column
            const int yk = y - lane;
            const bool valid = yk >= 1 && x > lb - yk;
            uint32_t op = BP_D, len = 1;
            if (valid) {
                if (layout == 2) {
                    const int sk = yk + x / C;
                    const size_t widx = (size_t)((sk - 1) >> 2) * J.stride + (size_t)((x / C) * CP + (x % C));
                    const uint32_t b = (w32[widx] >> (8 * (3 - ((sk - 1) & 3)))) & 0xFFu;
                    len = (b & 63u) + 1u;
                    if ((b >> 6) == 1u) op = BP_D;
                    else if ((b >> 6) == 2u) op = BP_I;
                    else {
                        const int qc = codes[bwd ? J.qIdx - (uint32_t)(yk - 1) : J.qIdx + (uint32_t)(yk - 1)];
                        const int ri = yk - lb - 1 + x;
                        const int rcode = nib(bases, bwd ? J.rOff - (uint32_t)ri : J.rOff + (uint32_t)ri);
                        op = (qc == rcode) ? BP_M : BP_R;
                    }
                } else {
                    const size_t row = layout ? (size_t)(yk + x / C) : (size_t)yk;
                    const uint32_t c = mytb[row * J.stride + x];
                    op = c >> 14; len = c & 0x3FFFu;
                }
            }
            const bool diag = valid && op <= BP_R;
            const unsigned dmask = __ballot_sync(full, diag);
            if (!(dmask & 1u)) {                                         // a gap cell: its whole run in one step
                const uint32_t op0 = __shfl_sync(full, op, 0), len0 = __shfl_sync(full, len, 0);
                emit((int)op0, len0);
                if (op0 == BP_D) x -= (int)len0; else { y -= (int)len0; x += (int)len0; }
                continue;
            }
            const int p = (dmask == full) ? 32 : (__ffs((int)~dmask) - 1);                 // leading match/replace steps
            unsigned rmask = __ballot_sync(full, diag && op == BP_R);
            if (p < 32) rmask &= (1u << p) - 1u;
            for (int done = 0; done < p;) {                              // cut the mismatch mask into runs
                const unsigned bit = (rmask >> done) & 1u;
                const unsigned rest = (bit ? ~rmask : rmask) >> done;    // first position where the type changes
                int r = rest ? (__ffs((int)rest) - 1) : 32;
                if (r > p - done) r = p - done;
                emit(bit ? (int)BP_R : (int)BP_M, (uint32_t)r);
                done += r;
            }
            y -= p;
            guard -= p - 1;
        }
        if (prev >= 0) { put(n, prev, run); n++; }
    }
    if (lane == 0) { o.n_ops = n; outs[t] = o; }
    // result record, fused (finalize_kernel of the thread-per-job path); the run count goes to the scan that
    // places every job's runs in job order
    if (lane == 0) {
        ya_dp_result r;
        uint32_t keep = (n >= 0xFFFFFFF0u) ? 0u : n;                     // traceback error marker: reported by the host
        if (inPlace) {
            const unsigned long long cells = ((unsigned long long)o.cells_hi << 32) | o.cells_lo;
            atomicAdd(&acct[0], cells);
            if (J.layout == 2) atomicAdd(&acct[1], cells);
            if (n >= 0xFFFFFFF0u) atomicOr(&acct[2], 1ull);
            else if (n > J.ops_cap) { atomicOr(&acct[2], 2ull); keep = 0; }
        }
        if (ext) {
            if (o.score <= 0) { r.score = 0; r.addedQLen = 0; r.addedRLen = 0; keep = 0; }                  // SW.cpp:525,1102
            else { r.score = o.score; r.addedQLen = (uint16_t)o.maxi; r.addedRLen = (uint16_t)(o.maxi + (o.maxj - 2 * BW)); }   // SW.cpp:1109-1110
        } else { r.score = o.score; r.addedQLen = 0; r.addedRLen = 0; }
        r.ops_n = keep; r.ops_off = 0;
        if (inPlace) r.ops_off = J.ops_off + ((J.kind != YA_DP_EXT_BWD) ? J.ops_cap - keep : 0u);
        res[t] = r;
        if (ops_cnt) ops_cnt[t] = keep;
    }
}

__global__ void finalize_kernel(const DevJob *__restrict__ jobs, const DevJobOut *__restrict__ outs, int n_jobs,
                                int BW, ya_dp_result *__restrict__ res, uint32_t *__restrict__ ops_cnt)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    const DevJob J = jobs[t];
    DevJobOut o = outs[t];
    if (o.n_ops >= 0xFFFFFFF0u) o.n_ops = 0;          // traceback error marker, reported by the host
    ya_dp_result r;
    r.ops_off = 0;
    if (J.kind >= YA_DP_EXT_FWD) {
        if (o.score <= 0) { r.score = 0; r.addedQLen = 0; r.addedRLen = 0; r.ops_n = 0; }       // SW.cpp:525,1102
        else {
            r.score = o.score;
            r.addedQLen = (uint16_t)o.maxi;                                                      // SW.cpp:1109
            r.addedRLen = (uint16_t)(o.maxi + (o.maxj - 2 * BW));                                // SW.cpp:1110
            r.ops_n = o.n_ops;
        }
    } else { r.score = o.score; r.addedQLen = 0; r.addedRLen = 0; r.ops_n = o.n_ops; }
    res[t] = r;
    ops_cnt[t] = r.ops_n;
}

// Copies each job's runs to their compact position.  Walking order is end -> start; forward and
// global jobs are reversed into genome order, backward extensions already are (SW.cpp:1184-1185).
// (runs that would not fit ops_out_cap are not written: the host sees total > capacity and runs the kernel again
//  on a larger array)
__global__ void compact_ops_kernel(const DevJob *__restrict__ jobs, int n_jobs, const uint32_t *__restrict__ ops_off,
                                   ya_dp_result *__restrict__ res, const ya_op *__restrict__ ops_raw,
                                   ya_op *__restrict__ ops_out, uint32_t ops_out_cap)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    const DevJob J = jobs[t];
    uint32_t n = res[t].ops_n, o = ops_off[t];
    res[t].ops_off = o;
    if ((uint64_t)o + n > (uint64_t)ops_out_cap) return;
    const ya_op *src = ops_raw + J.ops_off;
    const bool keepOrder = J.kind == YA_DP_EXT_BWD;
    for (uint32_t k = 0; k < n; k++) ops_out[o + k] = keepOrder ? src[k] : src[n - 1 - k];
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct WaveCfg { int G, C; };
static const WaveCfg kWaveCfgs[] = {{8, 3}, {8, 4}, {8, 6}, {8, 8}, {16, 6}, {16, 8}, {32, 8}, {32, 11}};
static const int kNumWaveCfgs = sizeof(kWaveCfgs) / sizeof(kWaveCfgs[0]);

template <int G, int C>
static void launch_wave(ya_ctx *c, bool ext, const uint32_t *d_ids, int n, const DpConst &K)
{
    const int threads = 128;
    const int groups_per_block = threads / G;
    int blocks = (n + groups_per_block - 1) / groups_per_block;
    if (ext)
        dp_wave_kernel<G, C, true><<<blocks, threads, 0, c->stream>>>(c->d_jobs.as<DevJob>(), d_ids, n, c->d_jobout.as<DevJobOut>(),
            c->d_tb.as<uint16_t>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), K);
    else
        dp_wave_kernel<G, C, false><<<blocks, threads, 0, c->stream>>>(c->d_jobs.as<DevJob>(), d_ids, n, c->d_jobout.as<DevJobOut>(),
            c->d_tb.as<uint16_t>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), K);
    c->ctr.launches++;
}

static void launch_wave_cfg(ya_ctx *c, int cfg, bool ext, const uint32_t *d_ids, int n, const DpConst &K)
{
    switch (cfg) {
    case 0: launch_wave<8, 3>(c, ext, d_ids, n, K); break;
    case 1: launch_wave<8, 4>(c, ext, d_ids, n, K); break;
    case 2: launch_wave<8, 6>(c, ext, d_ids, n, K); break;
    case 3: launch_wave<8, 8>(c, ext, d_ids, n, K); break;
    case 4: launch_wave<16, 6>(c, ext, d_ids, n, K); break;
    case 5: launch_wave<16, 8>(c, ext, d_ids, n, K); break;
    case 6: launch_wave<32, 8>(c, ext, d_ids, n, K); break;
    default: launch_wave<32, 11>(c, ext, d_ids, n, K); break;
    }
}

// YA_DP_MODE environment switch (tests): "thread" forces the generic kernel for every job,
// "wave" keeps extensions on dp_wave_kernel instead of the packed kernel.
static bool force_thread_kernel()
{
    const char *e = getenv("YA_DP_MODE");
    return e && strcmp(e, "thread") == 0;
}
static bool forbid_packed_kernel()
{
    const char *e = getenv("YA_DP_MODE");
    return e && (strcmp(e, "thread") == 0 || strcmp(e, "wave") == 0);
}

struct PackedCfg { int G, C, W; };
static const PackedCfg kPackedCfgs[] = {{2, 11, 21}, {4, 11, 41}};
// Below this many jobs per launch the grid cannot fill the GPU with 8 (16) jobs per warp; the narrow
// variants <7,6,41> / <4,6,21> put 4 (8) jobs in a warp: twice the warps, shorter macro steps.
// (level 0: <2,11,21> / <4,11,41>; level 1 below YA_PACKED_NARROW_BELOW jobs: <4,6,21> / <7,6,41>; level 2 below
//  YA_PACKED_XNARROW_BELOW: <7,3,21> / <14,3,41>, 4 / 2 jobs per warp -- a launch needs about 8 resident warps per scheduler
//  before its issue slots rather than one job's dependent chain set the pace)
static int packed_level(size_t nExt)
{
    static const size_t narrowBelow = [] { const char *e = getenv("YA_PACKED_NARROW_BELOW"); return e ? (size_t)atol(e) : (size_t)28000; }();
    static const size_t xnarrowBelow = [] { const char *e = getenv("YA_PACKED_XNARROW_BELOW"); return e ? (size_t)atol(e) : (size_t)5000; }();
    return nExt < xnarrowBelow ? 2 : nExt < narrowBelow ? 1 : 0;
}
static inline void packed_geom(int level, int pcls, int &G, int &C)
{
    static const int g[3][2] = {{2, 4}, {4, 7}, {7, 14}}, cc[3] = {11, 6, 3};
    G = g[level][pcls]; C = cc[level];
}
template <int G, int C, int W> static void launch_packed(ya_ctx *c, const uint32_t *d_ids, int n, const DpConst &K);
static void launch_packed_level(ya_ctx *c, int level, int pcls, const uint32_t *d_ids, int n, const DpConst &K);
static const int kNumPackedCfgs = sizeof(kPackedCfgs) / sizeof(kPackedCfgs[0]);

// YA_EXT_STAGE=1: the reference windows of the extension kernel staged into shared memory by bulk copies (see above).
static bool ext_stage_on()
{
    static const bool on = [] { const char *e = getenv("YA_EXT_STAGE"); return e && atoi(e) != 0; }();
    return on;
}
template <int G, int C, int W>
static void launch_packed(ya_ctx *c, const uint32_t *d_ids, int n, const DpConst &K)
{
    const int threads = 128;
    const int groups_per_block = (threads / 32) * (32 / G);
    int blocks = (n + groups_per_block - 1) / groups_per_block;
    if (ext_stage_on())
        dp_ext_packed_kernel<G, C, W, true><<<blocks, threads, 0, c->stream>>>(c->d_jobs.as<DevJob>(), d_ids, n, c->d_jobout.as<DevJobOut>(),
            c->d_tb.as<uint32_t>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), K);
    else
        dp_ext_packed_kernel<G, C, W><<<blocks, threads, 0, c->stream>>>(c->d_jobs.as<DevJob>(), d_ids, n, c->d_jobout.as<DevJobOut>(),
            c->d_tb.as<uint32_t>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), K);
    c->ctr.launches++;
}

static void launch_packed_level(ya_ctx *c, int level, int pcls, const uint32_t *d_ids, int n, const DpConst &K)
{
    if (pcls == 0) {
        if (level == 0) launch_packed<2, 11, 21>(c, d_ids, n, K); else if (level == 1) launch_packed<4, 6, 21>(c, d_ids, n, K); else launch_packed<7, 3, 21>(c, d_ids, n, K);
    } else {
        if (level == 0) launch_packed<4, 11, 41>(c, d_ids, n, K); else if (level == 1) launch_packed<7, 6, 41>(c, d_ids, n, K); else launch_packed<14, 3, 41>(c, d_ids, n, K);
    }
}

#include <chrono>
static double g_prof_sw[6];
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct SwProfPrinter { ~SwProfPrinter() { if (getenv("YA_PROF")) fprintf(stderr, "ya_sw_batch wall s: prep %.4f alloc+h2d %.4f launch..scan-sync %.4f compact+d2h-sync %.4f post %.4f\n",
    g_prof_sw[0], g_prof_sw[1], g_prof_sw[2], g_prof_sw[3], g_prof_sw[4]); } } g_sw_prof_printer;

extern "C" int ya_sw_batch(ya_ctx *c, const ya_dp_job *jobs, int n, ya_dp_result *res,
                           ya_op *ops, size_t ops_cap, size_t *ops_needed)
{
    double tp0 = now_s();
    if (!c || n < 0 || (n && (!jobs || !res))) return YA_E_ARG;
    if (ops_needed) *ops_needed = 0;
    c->ops_pending = 0;
    if (n == 0) return YA_OK;
    if (c->n_reads == 0) return ya_fail(c, YA_E_STATE, "ya_sw_batch: no read batch uploaded");
    YA_CUDA(c, cudaSetDevice(c->device));
    // bulk calls go to the low-priority stream (unless the caller installed its own stream)
    struct StreamSwap {
        ya_ctx *c; cudaStream_t saved;
        StreamSwap(ya_ctx *c_, bool bulk) : c(c_), saved(c_->stream) { if (bulk && c->stream == c->own_stream && c->bulk_stream) c->stream = c->bulk_stream; }
        ~StreamSwap() { c->stream = saved; }
    } streamSwap(c, n >= 2048);
    cudaStream_t st = c->stream;
    AllocScope allocScope(st);
    const ya_params &P = c->P;
    const int bw2 = 2 * P.bandWidth;
    const bool forceThread = force_thread_kernel();
    static const int fullThreadMaxW = [] { const char *e = getenv("YA_FULL_THREAD_MAXW"); return e ? atoi(e) : 0; }();   // full-matrix jobs up to this band width stay one thread per job
    const bool allowPacked = !forbid_packed_kernel();
    const int64_t packedStepCost = std::max<int64_t>(std::max<int64_t>(std::abs(P.MScore), std::abs(P.RCost)),
                                                     (int64_t)std::abs(P.GOCost) + std::abs(P.GECost));

    YA_CUDA(c, c->h_jobs.reserve((size_t)n * sizeof(DevJob)));
    DevJob *hj = c->h_jobs.as<DevJob>();
    // job id lists per kernel class: [0..kNumWaveCfgs) ext, [kNumWaveCfgs..2k) banded, last = thread
    // (persistent per-context vectors: no allocation on the steady-state path)
    const int nClasses = 2 * kNumWaveCfgs + 1 + kNumPackedCfgs;
    std::vector<std::vector<uint32_t>> &lists = c->sw_lists;
    if ((int)lists.size() != nClasses) lists.assign((size_t)nClasses, std::vector<uint32_t>());
    for (auto &v : lists) v.clear();
    const int packedBase = 2 * kNumWaveCfgs + 1;
    size_t nExt = 0;
    for (int i = 0; i < n; i++) nExt += jobs[i].kind >= YA_DP_EXT_FWD;
    const int packedLevel = packed_level(nExt);
    uint64_t tb_cells = 0, rows_ints = 0, ops_slots = 0;
    int n_live = 0;
    std::vector<uint32_t> &live_of = c->sw_live_of;   // device job index -> caller job index
    if ((int)live_of.size() < n) live_of.resize((size_t)n);
    for (int i = 0; i < n; i++) {
        const ya_dp_job &j = jobs[i];
        res[i].score = 0; res[i].addedQLen = 0; res[i].addedRLen = 0; res[i].ops_off = 0; res[i].ops_n = 0;
        if (j.read >= (uint32_t)c->n_reads || j.kind > YA_DP_EXT_BWD || j.strand > 1)
            return ya_fail(c, YA_E_ARG, "ya_sw_batch: bad job");
        const uint64_t rbase = c->h_read_off[j.read];
        const int L = (int)(c->h_read_off[j.read + 1] - rbase);
        int qLen = j.qLen, rLen = j.rLen;
        uint32_t rOff = j.rOff;
        DevJob d{};
        d.kind = j.kind; d.strand = j.strand;
        if (j.kind >= YA_DP_EXT_FWD) {
            // clamping of findAGSExtension (SW.cpp:492-516)
            const bool reverse = j.kind == YA_DP_EXT_BWD;
            if (qLen <= 0) continue;
            uint32_t rl = (uint32_t)(qLen + bw2);
            if (reverse && rl > rOff) { rl = rOff + 1; qLen = (int)(rl - (uint32_t)bw2); if (qLen <= 0) continue; }
            if (!reverse && rOff + rl > c->maxROff) { rl = c->maxROff - rOff; qLen = (int)(rl - (uint32_t)bw2); if (qLen <= 0) continue; }
            rLen = (int)rl;
            if (reverse ? ((int)j.qOff - (qLen - 1) < 0 || (int)j.qOff >= L) : ((int)j.qOff + qLen > L))
                return ya_fail(c, YA_E_ARG, "ya_sw_batch: extension runs outside the read");
            d.lb = (uint16_t)bw2; d.rb = (uint16_t)bw2;
        } else {
            if (qLen <= 0 || rLen <= 0) return ya_fail(c, YA_E_ARG, "ya_sw_batch: global job with empty side");
            if ((int)j.qOff + qLen > L) return ya_fail(c, YA_E_ARG, "ya_sw_batch: global job runs outside the read");
            if ((uint64_t)rOff + (uint64_t)rLen > (uint64_t)c->n_base_bytes * 2)
                return ya_fail(c, YA_E_ARG, "ya_sw_batch: global job runs outside the reference");
            if (j.kind == YA_DP_BANDED) {
                d.lb = (uint16_t)(P.bandWidth + (qLen > rLen ? qLen - rLen : 0));     // SW.cpp:856-866
                d.rb = (uint16_t)(P.bandWidth + (rLen > qLen ? rLen - qLen : 0));
            } else { d.lb = (uint16_t)qLen; d.rb = (uint16_t)rLen; }                  // full matrix as a band
        }
        d.rOff = rOff; d.rLen = (uint16_t)rLen; d.qLen = (uint16_t)qLen;
        d.qIdx = (uint32_t)(rbase + j.qOff);
        const int W = d.lb + d.rb + 1;
        int cls = -1, pcls = -1;
        // The packed kernel keeps scores x256 in int32 below a sentinel of -2^29: a job is eligible only while every real
        // score (|V| <= (rows + W + 2) * largest per-step cost) and the drift of the sentinel chains (one cost per row)
        // stay below 2^28 / 256 = 2^20; anything larger runs on dp_wave_kernel, which computes in plain int32 like SW.cpp.
        if (allowPacked && j.kind >= YA_DP_EXT_FWD && W <= P.maxGap && W <= P.maxIntron &&
            (int64_t)(qLen + W + 2) * packedStepCost < ((int64_t)1 << 20))
            for (int k = 0; k < kNumPackedCfgs; k++) if (kPackedCfgs[k].W == W) pcls = k;
        if (pcls >= 0) {
            int G, C; packed_geom(packedLevel, pcls, G, C);
            const int CP = (C + 3) & ~3;
            d.layout = 2; d.colsPerLane = (uint8_t)C; d.stride = (uint32_t)(G * CP);
            tb_cells = (tb_cells + 7) & ~7ull;                         // 16-byte alignment for the 128-bit stores
            d.tb_off = tb_cells / 2;                                  // in 32-bit words
            tb_cells += 2ull * (uint64_t)((qLen + G + 3) / 4 + 1) * d.stride;
            lists[packedBase + pcls].push_back((uint32_t)n_live);
        } else if (!forceThread && (j.kind != YA_DP_FULL || W > fullThreadMaxW)) {
            // (full-matrix jobs are bands with lb = qLen, rb = rLen: the small ones stay one thread per job,
            //  the larger ones take a lane group like a banded job -- their serial row chain is what a round waits for)
            for (int k = 0; k < kNumWaveCfgs; k++)
                if (kWaveCfgs[k].G * kWaveCfgs[k].C >= W) { cls = k; break; }
        }
        if (pcls >= 0) {
        } else if (cls >= 0) {
            const int G = kWaveCfgs[cls].G, C = kWaveCfgs[cls].C;
            d.layout = 1; d.colsPerLane = (uint8_t)C; d.stride = (uint32_t)(G * C);
            d.tb_off = tb_cells;
            tb_cells += (uint64_t)(qLen + G + 1) * d.stride;
            lists[(j.kind >= YA_DP_EXT_FWD ? 0 : kNumWaveCfgs) + cls].push_back((uint32_t)n_live);
        } else {
            d.layout = 0; d.colsPerLane = 1; d.stride = (uint32_t)W;
            d.tb_off = tb_cells;
            tb_cells += (uint64_t)(qLen + 1) * d.stride;
            d.rows_off = (uint32_t)rows_ints;
            rows_ints += 3ull * (W + 1);
            lists[2 * kNumWaveCfgs].push_back((uint32_t)n_live);
        }
        d.ops_off = (uint32_t)ops_slots;
        d.ops_cap = (uint32_t)(qLen + rLen + 2);
        ops_slots += d.ops_cap;
        if (ops_slots >= 0xFFFF0000ull || rows_ints >= 0xFFFF0000ull)
            return ya_fail(c, YA_E_ARG, "ya_sw_batch: batch too large, split it");
        hj[n_live] = d;
        live_of[n_live] = (uint32_t)i;
        n_live++;
    }
    c->ctr.dp_jobs += (uint64_t)n;
    if (n_live == 0) return YA_OK;
    double tp1 = now_s(); g_prof_sw[0] += tp1 - tp0;

    DeviceTurn turn(c->device);
    YA_CUDA(c, c->d_jobs.reserve((size_t)n_live * sizeof(DevJob)));
    YA_CUDA(c, c->d_jobout.reserve((size_t)n_live * sizeof(DevJobOut)));
    YA_CUDA(c, c->d_tb.reserve(tb_cells * 2 + 64));
    YA_CUDA(c, c->d_rows.reserve(rows_ints * 4 + 64));
    YA_CUDA(c, c->d_ops_raw.reserve(ops_slots * sizeof(ya_op) + 64));
    YA_CUDA(c, c->d_ops_cnt.reserve((size_t)n_live * 4 + 64));
    YA_CUDA(c, c->d_ops_off.reserve((size_t)n_live * 4 + 64));
    YA_CUDA(c, c->d_res.reserve((size_t)n_live * sizeof(ya_dp_result)));
    YA_CUDA(c, c->d_misc.reserve((size_t)n_live * 4 + 64));
    YA_CUDA(c, c->h_res.reserve((size_t)n_live * (sizeof(ya_dp_result) + sizeof(DevJobOut)) + 64));
    YA_CUDA(c, cudaMemcpyAsync(c->d_jobs.p, hj, (size_t)n_live * sizeof(DevJob), cudaMemcpyHostToDevice, st));
    // longest jobs first inside each kernel class (groups / lanes sharing a warp get similar row counts,
    // in the fill kernels and in the traceback): counting sort on qLen/8, descending -- a comparison
    // sort of 20 K ids costs milliseconds here
    for (size_t k = 0; k < lists.size(); k++) {
        std::vector<uint32_t> &v = lists[k];
        if (v.size() < 64) continue;
        std::vector<uint32_t> &cnt = c->sw_cnt, &tmp = c->sw_tmp;
        cnt.assign(8192 + 1, 0u);
        for (uint32_t id : v) cnt[8191 - (hj[id].qLen >> 3)]++;
        uint32_t run = 0;
        for (size_t b = 0; b <= 8192; b++) { uint32_t t = cnt[b]; cnt[b] = run; run += t; }
        tmp.resize(v.size());
        for (uint32_t id : v) tmp[cnt[8191 - (hj[id].qLen >> 3)]++] = id;
        v.swap(tmp);
    }
    // id lists
    std::vector<uint32_t> &flat = c->sw_flat; flat.clear(); flat.reserve(n_live);
    std::vector<size_t> start(lists.size());
    for (size_t k = 0; k < lists.size(); k++) { start[k] = flat.size(); flat.insert(flat.end(), lists[k].begin(), lists[k].end()); }
    uint32_t *d_ids = c->d_misc.as<uint32_t>();
    YA_CUDA(c, cudaMemcpyAsync(d_ids, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, st));

    DpConst K{P.GOCost, P.GECost, P.RCost, P.MScore, P.XCutoff, P.maxIntron, P.maxGap, P.bandWidth};
    double tp2 = now_s(); g_prof_sw[1] += tp2 - tp1;
    YA_CUDA(c, cudaEventRecord(c->ev[0], st));
    for (int k = 0; k < kNumWaveCfgs; k++) {
        if (!lists[k].empty()) launch_wave_cfg(c, k, true, d_ids + start[k], (int)lists[k].size(), K);
        if (!lists[kNumWaveCfgs + k].empty())
            launch_wave_cfg(c, k, false, d_ids + start[kNumWaveCfgs + k], (int)lists[kNumWaveCfgs + k].size(), K);
    }
    const bool anyPacked = !lists[packedBase + 0].empty() || !lists[packedBase + 1].empty();
    // YA_BULK_EXCLUSIVE=1: bulk extension launches of the contexts sharing a device run one at a time, which keeps
    // their CUDA-event timing free of another pipeline's bulk kernel (costs one more wait per bulk call).
    static const bool bulkExclusive = [] { const char *e = getenv("YA_BULK_EXCLUSIVE"); return e && atoi(e) != 0; }();   // off by default
    const size_t nPackedAll = lists[packedBase + 0].size() + lists[packedBase + 1].size();
    std::unique_lock<std::mutex> bulkTurn(ya_bulk_mutex(c->device), std::defer_lock);
    if (anyPacked && bulkExclusive && nPackedAll >= 2048) bulkTurn.lock();
    if (anyPacked) YA_CUDA(c, cudaEventRecord(c->ev[3], st));
    if (!lists[packedBase + 0].empty()) {
        launch_packed_level(c, packedLevel, 0, d_ids + start[packedBase + 0], (int)lists[packedBase + 0].size(), K);
    }
    if (!lists[packedBase + 1].empty()) {
        launch_packed_level(c, packedLevel, 1, d_ids + start[packedBase + 1], (int)lists[packedBase + 1].size(), K);
    }
    if (anyPacked) YA_CUDA(c, cudaEventRecord(c->ev[4], st));
    if (bulkTurn.owns_lock()) { YA_CUDA(c, ya_event_wait(c->ev[4])); bulkTurn.unlock(); }
    if (!lists[2 * kNumWaveCfgs].empty()) {
        int nt = (int)lists[2 * kNumWaveCfgs].size();
        dp_thread_kernel<<<(nt + 63) / 64, 64, 0, st>>>(c->d_jobs.as<DevJob>(), d_ids + start[2 * kNumWaveCfgs], nt,
            c->d_jobout.as<DevJobOut>(), c->d_tb.as<uint16_t>(), c->d_rows.as<int>(), c->d_bases,
            c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), K);
        c->ctr.launches++;
    }
    YA_CUDA(c, cudaEventRecord(c->ev[1], st));
    const int tbk = (n_live + 127) / 128;
    static const bool tbThread = [] { const char *e = getenv("YA_TB"); return e && strcmp(e, "thread") == 0; }();
    uint32_t *d_tot = c->d_ops_cnt.as<uint32_t>() + n_live;      // spare word after the counts
    ya_dp_result *hres = c->h_res.as<ya_dp_result>();
    DevJobOut *hout = (DevJobOut *)(hres + n_live);
    uint32_t *h_total = (uint32_t *)(hout + n_live);              // (page-locked: a pageable target would stage the copy)
    uint32_t total_ops = 0;
    bool fits = false;
    // the three-kernel epilogue (result records, scan of the run counts, compaction) with a sync in the middle for the total
    auto epilogueKernels = [&]() -> int {
        finalize_kernel<<<tbk, 128, 0, st>>>(c->d_jobs.as<DevJob>(), c->d_jobout.as<DevJobOut>(), n_live, P.bandWidth,
                                             c->d_res.as<ya_dp_result>(), c->d_ops_cnt.as<uint32_t>());
        c->ctr.launches++;
        int rc = ya_exclusive_scan_u32(c, c->d_ops_cnt.as<uint32_t>(), c->d_ops_off.as<uint32_t>(), (size_t)n_live, d_tot);
        if (rc != YA_OK) return rc;
        YA_CUDA(c, cudaMemcpyAsync(h_total, d_tot, 4, cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, ya_stream_wait(st));
        total_ops = *h_total;
        YA_CUDA(c, c->d_ops_out.reserve((size_t)total_ops * sizeof(ya_op) + 64));
        compact_ops_kernel<<<tbk, 128, 0, st>>>(c->d_jobs.as<DevJob>(), n_live, c->d_ops_off.as<uint32_t>(),
                                                c->d_res.as<ya_dp_result>(), c->d_ops_raw.as<ya_op>(), c->d_ops_out.as<ya_op>(),
                                                (uint32_t)std::min<size_t>(c->d_ops_out.cap / sizeof(ya_op), 0xFFFFFFFFu));
        c->ctr.launches++;
        return YA_OK;
    };
    double tp3 = tp2;
    if (tbThread) {
        traceback_kernel<<<tbk, 128, 0, st>>>(c->d_jobs.as<DevJob>(), d_ids, n_live, c->d_jobout.as<DevJobOut>(),
                                              c->d_tb.as<uint16_t>(), c->d_ops_raw.as<ya_op>(), c->d_bases,
                                              c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>());
        c->ctr.launches++;
        int rc = epilogueKernels();
        if (rc != YA_OK) return rc;
        tp3 = now_s(); g_prof_sw[2] += tp3 - tp2;
        YA_CUDA(c, cudaEventRecord(c->ev[2], st));
        YA_CUDA(c, cudaMemcpyAsync(hres, c->d_res.p, (size_t)n_live * sizeof(ya_dp_result), cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, cudaMemcpyAsync(hout, c->d_jobout.p, (size_t)n_live * sizeof(DevJobOut), cudaMemcpyDeviceToHost, st));
        fits = total_ops <= ops_cap && (total_ops == 0 || ops != nullptr);
        if (fits && total_ops)
            YA_CUDA(c, cudaMemcpyAsync(ops, c->d_ops_out.p, (size_t)total_ops * sizeof(ya_op), cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, ya_stream_wait(st));
    } else {
        // warp-per-job traceback with the result records fused in, then scan + compaction without a host round
        // trip: ONE synchronisation per call.  The compact run array is sized for 32 runs per job up front and the
        // first 24 per job are copied back speculatively together with the results.
        const size_t capDev = std::max<size_t>((size_t)32 * (size_t)n_live + 1024, c->d_ops_out.cap / sizeof(ya_op));
        YA_CUDA(c, c->d_ops_out.reserve(capDev * sizeof(ya_op)));
        const size_t devCap = std::min<size_t>(c->d_ops_out.cap / sizeof(ya_op), 0xFFFFFFFFu);
        traceback_warp_kernel<<<(n_live + 3) / 4, 128, 0, st>>>(c->d_jobs.as<DevJob>(), d_ids, n_live, c->d_jobout.as<DevJobOut>(),
                                                                 c->d_tb.as<uint16_t>(), c->d_ops_raw.as<ya_op>(), c->d_bases,
                                                                 c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(),
                                                                 P.bandWidth, c->d_res.as<ya_dp_result>(), c->d_ops_cnt.as<uint32_t>(), nullptr);
        c->ctr.launches++;
        {
            int rc = ya_exclusive_scan_u32(c, c->d_ops_cnt.as<uint32_t>(), c->d_ops_off.as<uint32_t>(), (size_t)n_live, d_tot);
            if (rc != YA_OK) return rc;
        }
        compact_ops_kernel<<<tbk, 128, 0, st>>>(c->d_jobs.as<DevJob>(), n_live, c->d_ops_off.as<uint32_t>(),
                                                c->d_res.as<ya_dp_result>(), c->d_ops_raw.as<ya_op>(), c->d_ops_out.as<ya_op>(),
                                                (uint32_t)devCap);
        c->ctr.launches++;
        YA_CUDA(c, cudaEventRecord(c->ev[2], st));
        YA_CUDA(c, cudaMemcpyAsync(h_total, d_tot, 4, cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, cudaMemcpyAsync(hres, c->d_res.p, (size_t)n_live * sizeof(ya_dp_result), cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, cudaMemcpyAsync(hout, c->d_jobout.p, (size_t)n_live * sizeof(DevJobOut), cudaMemcpyDeviceToHost, st));
        const size_t guess = ops ? std::min<size_t>(std::min<size_t>((size_t)24 * (size_t)n_live + 256, devCap), ops_cap) : 0;
        if (guess) YA_CUDA(c, cudaMemcpyAsync(ops, c->d_ops_out.p, guess * sizeof(ya_op), cudaMemcpyDeviceToHost, st));
        YA_CUDA(c, ya_stream_wait(st));
        tp3 = now_s(); g_prof_sw[2] += tp3 - tp2;
        total_ops = *h_total;
        if ((size_t)total_ops > devCap) {
            // more runs than the compact array holds (rare at 32 per job): compact again into a larger array
            YA_CUDA(c, c->d_ops_out.reserve((size_t)total_ops * sizeof(ya_op) + 64));
            compact_ops_kernel<<<tbk, 128, 0, st>>>(c->d_jobs.as<DevJob>(), n_live, c->d_ops_off.as<uint32_t>(),
                                                    c->d_res.as<ya_dp_result>(), c->d_ops_raw.as<ya_op>(), c->d_ops_out.as<ya_op>(),
                                                    (uint32_t)std::min<size_t>(c->d_ops_out.cap / sizeof(ya_op), 0xFFFFFFFFu));
            c->ctr.launches++;
            YA_CUDA(c, cudaMemcpyAsync(hres, c->d_res.p, (size_t)n_live * sizeof(ya_dp_result), cudaMemcpyDeviceToHost, st));
            fits = total_ops <= ops_cap && (total_ops == 0 || ops != nullptr);
            if (fits && total_ops)
                YA_CUDA(c, cudaMemcpyAsync(ops, c->d_ops_out.p, (size_t)total_ops * sizeof(ya_op), cudaMemcpyDeviceToHost, st));
            YA_CUDA(c, ya_stream_wait(st));
        } else {
            fits = total_ops <= ops_cap && (total_ops == 0 || ops != nullptr);
            if (fits && (size_t)total_ops > guess) {                  // the speculative copy was too short: fetch the rest
                YA_CUDA(c, cudaMemcpyAsync(ops + guess, c->d_ops_out.as<ya_op>() + guess, ((size_t)total_ops - guess) * sizeof(ya_op),
                                           cudaMemcpyDeviceToHost, st));
                YA_CUDA(c, ya_stream_wait(st));
            }
        }
    }
    YA_CUDA(c, cudaGetLastError());
    turn.done();
    double tp4 = now_s(); g_prof_sw[3] += tp4 - tp3;
    float ms0 = 0, ms1 = 0;
    cudaEventElapsedTime(&ms0, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&ms1, c->ev[1], c->ev[2]);
    c->ctr.ms_dp += ms0; c->ctr.ms_traceback += ms1;
    // Roofline accounting covers the bulk launches only (>= 4096 jobs): the demand-driven careful
    // re-extension rounds launch a handful of jobs and are bound by one job's serial latency.
    const size_t nPackedJobs = lists[packedBase + 0].size() + lists[packedBase + 1].size();
    const bool bulkPacked = nPackedJobs >= 4096;
    if (bulkPacked) {
        float ms2 = 0;
        cudaEventElapsedTime(&ms2, c->ev[3], c->ev[4]);
        c->ctr.ms_ext += ms2;
        c->ctr.ext_launches += (lists[packedBase + 0].empty() ? 0 : 1) + (lists[packedBase + 1].empty() ? 0 : 1);
        ya_note_ext_interval(c, c->ev[3], c->ev[4]);
    }
    for (int k = 0; k < n_live; k++) {
        if (hout[k].n_ops >= 0xFFFFFFF0u) return ya_fail(c, YA_E_STATE, "internal: traceback walked off the band (corrupt back-pointers)");
        res[live_of[k]] = hres[k];
        const uint64_t jc = ((uint64_t)hout[k].cells_hi << 32) | hout[k].cells_lo;
        c->ctr.dp_cells += jc;
        if (hj[k].layout == 2 && bulkPacked) c->ctr.ext_cells += jc;
        if (hres[k].ops_n > hj[k].ops_cap) return ya_fail(c, YA_E_STATE, "internal: op scratch overflow");
    }
    if (ops_needed) *ops_needed = total_ops;
    g_prof_sw[4] += now_s() - tp4;
    if (!fits) { c->ops_pending = total_ops; return ya_fail(c, YA_E_CAPACITY, "op output buffer too small"); }
    return YA_OK;
}

extern "C" int ya_sw_fetch_ops(ya_ctx *c, ya_op *ops, size_t ops_cap)
{
    if (!c) return YA_E_ARG;
    if (c->ops_pending == 0) return ya_fail(c, YA_E_STATE, "ya_sw_fetch_ops: no edit operations pending");
    if (!ops || ops_cap < c->ops_pending) return ya_fail(c, YA_E_CAPACITY, "op output buffer too small");
    YA_CUDA(c, cudaSetDevice(c->device));
    YA_CUDA(c, cudaMemcpyAsync(ops, c->d_ops_out.p, c->ops_pending * sizeof(ya_op), cudaMemcpyDeviceToHost, c->stream));
    YA_CUDA(c, ya_stream_wait(c->stream));
    return YA_OK;
}

// ------------------------------------------------------------------------------------------
// A DP round whose jobs were born on the device (ya_prepare_clumps) and whose answers stay there (ya_align_batch): what
// the host loop of ya_sw_batch does per job -- clamping of findAGSExtension (SW.cpp:492-516), band geometry
// (SW.cpp:856-866), kernel class, back-pointer / run-slot layout, longest-first order inside a class -- as kernels.
// ------------------------------------------------------------------------------------------
#define DPR_NCLASS   (2 * 8 + 1 + 2)          // == 2 * kNumWaveCfgs + 1 + kNumPackedCfgs
#define DPR_DEAD     DPR_NCLASS               // jobs settled without a launch (nothing left after clamping)
#define DPR_BUCKETS  1024                     // rows / 8, capped: the order inside a class is by length, longest first
#define DPR_BINS     ((DPR_NCLASS + 1) * DPR_BUCKETS)

struct DprFlags { int packedLevel, allowPacked, forceThread, fullThreadMaxW; long long packedStepCost; };
__constant__ int c_waveG[8] = {8, 8, 8, 8, 16, 16, 32, 32};
__constant__ int c_waveC[8] = {3, 4, 6, 8, 6, 8, 8, 11};

__global__ void dpr_classify_kernel(const ya_dp_job *__restrict__ jobs, uint32_t n, const uint64_t *__restrict__ read_off, int n_reads,
                                    ya_params P, uint32_t maxROff, uint64_t n_bases, DprFlags F,
                                    DevJob *__restrict__ dj, uint32_t *__restrict__ key, uint32_t *__restrict__ tb_units,
                                    uint32_t *__restrict__ ops_slots, uint32_t *__restrict__ rows_ints,
                                    ya_dp_result *__restrict__ res, unsigned long long *__restrict__ acct)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ya_dp_job j = jobs[i];
    DevJob d; memset(&d, 0, sizeof d);
    ya_dp_result zero; zero.score = 0; zero.addedQLen = 0; zero.addedRLen = 0; zero.ops_off = 0; zero.ops_n = 0;
    res[i] = zero;
    uint32_t cls = DPR_DEAD, tbu = 0, slots = 0, rints = 0;
    const int bw2 = 2 * P.bandWidth;
    bool bad = j.read >= (uint32_t)n_reads || j.kind > YA_DP_EXT_BWD || j.strand > 1;
    if (!bad) {
        const uint64_t rbase = read_off[j.read];
        const int L = (int)(read_off[j.read + 1] - rbase);
        int qLen = j.qLen, rLen = j.rLen;
        const uint32_t rOff = j.rOff;
        bool live = true;
        d.kind = j.kind; d.strand = j.strand;
        if (j.kind >= YA_DP_EXT_FWD) {                                  // SW.cpp:492-516
            const bool reverse = j.kind == YA_DP_EXT_BWD;
            if (qLen <= 0) live = false;
            else {
                uint32_t rl = (uint32_t)(qLen + bw2);
                if (reverse && rl > rOff) { rl = rOff + 1; qLen = (int)(rl - (uint32_t)bw2); if (qLen <= 0) live = false; }
                if (live && !reverse && rOff + rl > maxROff) { rl = maxROff - rOff; qLen = (int)(rl - (uint32_t)bw2); if (qLen <= 0) live = false; }
                rLen = (int)rl;
                if (live && (reverse ? ((int)j.qOff - (qLen - 1) < 0 || (int)j.qOff >= L) : ((int)j.qOff + qLen > L))) bad = true;
                d.lb = (uint16_t)bw2; d.rb = (uint16_t)bw2;
            }
        } else {
            if (qLen <= 0 || rLen <= 0 || (int)j.qOff + qLen > L || (uint64_t)rOff + (uint64_t)rLen > n_bases) bad = true;
            else if (j.kind == YA_DP_BANDED) {
                d.lb = (uint16_t)(P.bandWidth + (qLen > rLen ? qLen - rLen : 0));     // SW.cpp:856-866
                d.rb = (uint16_t)(P.bandWidth + (rLen > qLen ? rLen - qLen : 0));
            } else { d.lb = (uint16_t)qLen; d.rb = (uint16_t)rLen; }
        }
        if (live && !bad) {
            d.rOff = rOff; d.rLen = (uint16_t)rLen; d.qLen = (uint16_t)qLen;
            d.qIdx = (uint32_t)(rbase + j.qOff);
            const int W = d.lb + d.rb + 1;
            int wcls = -1, pcls = -1;
            if (F.allowPacked && j.kind >= YA_DP_EXT_FWD && W <= P.maxGap && W <= P.maxIntron &&
                (long long)(qLen + W + 2) * F.packedStepCost < (1ll << 20)) {
                if (W == 21) pcls = 0; else if (W == 41) pcls = 1;
            }
            uint64_t cells;
            if (pcls >= 0) {
                const int G = F.packedLevel == 0 ? (pcls == 0 ? 2 : 4) : F.packedLevel == 1 ? (pcls == 0 ? 4 : 7) : (pcls == 0 ? 7 : 14);
                const int C = F.packedLevel == 0 ? 11 : F.packedLevel == 1 ? 6 : 3, CP = (C + 3) & ~3;
                d.layout = 2; d.colsPerLane = (uint8_t)C; d.stride = (uint32_t)(G * CP);
                cells = 2ull * (uint64_t)((qLen + G + 3) / 4 + 1) * d.stride;
                cls = 17 + pcls;
            } else {
                if (!F.forceThread && (j.kind != YA_DP_FULL || W > F.fullThreadMaxW))
                    for (int k = 0; k < 8; k++) if (c_waveG[k] * c_waveC[k] >= W) { wcls = k; break; }
                if (wcls >= 0) {
                    const int G = c_waveG[wcls], C = c_waveC[wcls];
                    d.layout = 1; d.colsPerLane = (uint8_t)C; d.stride = (uint32_t)(G * C);
                    cells = (uint64_t)(qLen + G + 1) * d.stride;
                    cls = (j.kind >= YA_DP_EXT_FWD ? 0 : 8) + wcls;
                } else {
                    d.layout = 0; d.colsPerLane = 1; d.stride = (uint32_t)W;
                    cells = (uint64_t)(qLen + 1) * d.stride;
                    rints = 3u * (uint32_t)(W + 1);
                    cls = 16;
                }
            }
            tbu = (uint32_t)((cells + 7) >> 3);                        // 16-byte units (128-bit stores of the packed kernel)
            slots = (uint32_t)(qLen + rLen + 2);
            d.ops_cap = slots;
        }
    }
    if (bad) atomicOr(&acct[2], 4ull);
    dj[i] = d;
    const uint32_t bucket = min((uint32_t)d.qLen >> 3, (uint32_t)(DPR_BUCKETS - 1));
    key[i] = cls * DPR_BUCKETS + (DPR_BUCKETS - 1 - bucket);
    tb_units[i] = tbu; ops_slots[i] = slots; rows_ints[i] = rints;
}

// In-place exclusive scans of up to four arrays of n words, ONE block per array (n is a few tens of thousands: a device-wide scan
// would be three launches per array); totals[k] receives array k's sum (64-bit).
__global__ void __launch_bounds__(1024)
dpr_scan_kernel(uint32_t *a0, uint32_t *a1, uint32_t *a2, uint32_t *a3, uint32_t n0, uint32_t n1, uint32_t n2, uint32_t n3,
                unsigned long long *__restrict__ totals)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    uint32_t *arr[4] = {a0, a1, a2, a3};
    const uint32_t cnt[4] = {n0, n1, n2, n3};
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int a = (int)blockIdx.x; a < 4; a += (int)gridDim.x) {             // (launched with four blocks: one array each)
        uint32_t *p = arr[a];
        if (!p) continue;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < cnt[a]; base += 1024 * 4) {
            const uint32_t i0 = base + threadIdx.x * 4;
            uint32_t v[4]; unsigned long long s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) { v[k] = (i0 + k < cnt[a]) ? p[i0 + k] : 0u; s += v[k]; }
            unsigned long long inc = s;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, dlt); if (lane >= dlt) inc += t; }
            if (lane == 31) wsum[w] = inc;
            __syncthreads();
            if (w == 0) {
                unsigned long long x = wsum[lane], xi = x;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, xi, dlt); if (lane >= dlt) xi += t; }
                wsum[lane] = xi - x;
            }
            __syncthreads();
            unsigned long long ex = carry + wsum[w] + inc - s;
#pragma unroll
            for (int k = 0; k < 4; k++) { if (i0 + k < cnt[a]) p[i0 + k] = (uint32_t)ex; ex += v[k]; }
            __syncthreads();
            if (threadIdx.x == 1023) carry = ex;
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[a] = carry;
        __syncthreads();
    }
}

__global__ void dpr_hist_kernel(const uint32_t *__restrict__ key, uint32_t n, uint32_t *__restrict__ bins)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&bins[key[i]], 1u);
}

// job offsets from the scans (the scanned arrays hold exclusive prefix sums now) and the class-ordered id list
__global__ void dpr_place_kernel(DevJob *__restrict__ dj, uint32_t n, const uint32_t *__restrict__ key, const uint32_t *__restrict__ tb_off,
                                 const uint32_t *__restrict__ ops_off, const uint32_t *__restrict__ rows_off,
                                 uint32_t *__restrict__ cursor, uint32_t *__restrict__ ids)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevJob d = dj[i];
    d.tb_off = (d.layout == 2) ? (uint64_t)tb_off[i] * 4 : (uint64_t)tb_off[i] * 8;    // packed: 32-bit words; else 16-bit cells
    d.ops_off = ops_off[i];
    d.rows_off = rows_off[i];
    dj[i] = d;
    ids[atomicAdd(&cursor[key[i]], 1u)] = i;
}

__global__ void dpr_class_starts_kernel(const uint32_t *__restrict__ bins_scanned, uint32_t *__restrict__ starts)
{
    const int k = threadIdx.x;
    if (k <= DPR_NCLASS + 1) starts[k] = (k <= DPR_NCLASS) ? bins_scanned[k * DPR_BUCKETS] : 0u;
}

// Runs the round for c->d_pc_jobs[0..n_jobs).  Afterwards c->d_res[i] answers job i with ops_off addressing c->d_ops_raw
// directly (runs in genome order); nothing is copied to the host.  acct (device, 4 x u64) accumulates cells / packed cells /
// error flags.  One synchronisation (the totals that size the scratch).
int ya_sw_device_round(ya_ctx *c, uint32_t n_jobs, uint32_t n_ext, unsigned long long *d_acct, size_t *raw_slots)
{
    if (raw_slots) *raw_slots = 0;
    c->dpr_ran = false;
    if (n_jobs == 0) return YA_OK;
    cudaStream_t st = c->stream;
    const ya_params &P = c->P;
    static const int fullThreadMaxW = [] { const char *e = getenv("YA_FULL_THREAD_MAXW"); return e ? atoi(e) : 0; }();
    DprFlags F;
    F.packedLevel = packed_level(n_ext); F.allowPacked = !forbid_packed_kernel(); F.forceThread = force_thread_kernel();
    F.fullThreadMaxW = fullThreadMaxW;
    F.packedStepCost = std::max<long long>(std::max<long long>(std::abs(P.MScore), std::abs(P.RCost)), (long long)std::abs(P.GOCost) + std::abs(P.GECost));
    YA_CUDA(c, c->d_jobs.reserve((size_t)n_jobs * sizeof(DevJob)));
    YA_CUDA(c, c->d_jobout.reserve((size_t)n_jobs * sizeof(DevJobOut)));
    YA_CUDA(c, c->d_res.reserve((size_t)n_jobs * sizeof(ya_dp_result)));
    YA_CUDA(c, c->d_misc.reserve((size_t)n_jobs * 4 + 64));
    YA_CUDA(c, c->d_dpr.reserve(8 * 8 + ((size_t)4 * n_jobs + DPR_BINS + 64) * 4));
    YA_CUDA(c, c->h_stage3.reserve(1024));
    unsigned long long *totals = c->d_dpr.as<unsigned long long>();                      // 4 totals (8 words reserved)
    uint32_t *key = (uint32_t *)(totals + 8), *tbu = key + n_jobs, *slots = tbu + n_jobs, *rints = slots + n_jobs;
    uint32_t *bins = rints + n_jobs, *starts = bins + DPR_BINS;                          // (starts: DPR_NCLASS + 2 words)
    uint32_t *d_ids = c->d_misc.as<uint32_t>();
    const unsigned nb = (n_jobs + 255) / 256;
    YA_CUDA(c, cudaMemsetAsync(bins, 0, (size_t)DPR_BINS * 4, st));
    dpr_classify_kernel<<<nb, 256, 0, st>>>(c->d_pc_jobs.as<ya_dp_job>(), n_jobs, c->d_read_off.as<uint64_t>(), c->n_reads, P, c->maxROff,
                                            (uint64_t)c->n_base_bytes * 2, F, c->d_jobs.as<DevJob>(), key, tbu, slots, rints,
                                            c->d_res.as<ya_dp_result>(), d_acct);
    dpr_hist_kernel<<<nb, 256, 0, st>>>(key, n_jobs, bins);
    dpr_scan_kernel<<<4, 1024, 0, st>>>(tbu, slots, rints, bins, n_jobs, n_jobs, n_jobs, DPR_BINS, totals);
    dpr_class_starts_kernel<<<1, 64, 0, st>>>(bins, starts);
    dpr_place_kernel<<<nb, 256, 0, st>>>(c->d_jobs.as<DevJob>(), n_jobs, key, tbu, slots, rints, bins, d_ids);
    c->ctr.launches += 5;
    struct HostPlan { unsigned long long totals[4]; uint32_t starts[DPR_NCLASS + 2]; } *hp = c->h_stage3.as<HostPlan>();
    YA_CUDA(c, cudaMemcpyAsync(hp->totals, totals, 4 * 8, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, cudaMemcpyAsync(hp->starts, starts, (DPR_NCLASS + 2) * 4, cudaMemcpyDeviceToHost, st));
    YA_CUDA(c, ya_stream_wait(st));
    const unsigned long long tbUnits = hp->totals[0], opsSlots = hp->totals[1], rowsInts = hp->totals[2];
    if (tbUnits >= 0xFFFF0000ull || opsSlots >= 0xFFFF0000ull || rowsInts >= 0xFFFF0000ull)
        return ya_fail(c, YA_E_STATE, "device DP round too large for 32-bit offsets");
    uint32_t start[DPR_NCLASS + 2];
    for (int k = 0; k <= DPR_NCLASS; k++) start[k] = hp->starts[k];
    start[DPR_NCLASS + 1] = n_jobs;
    auto count = [&](int k) { return (int)(start[k + 1] - start[k]); };
    const uint32_t n_live = start[DPR_NCLASS];                          // dead jobs sit behind every class
    c->ctr.dp_jobs += n_jobs;
    if (raw_slots) *raw_slots = (size_t)opsSlots;
    if (n_live == 0) return YA_OK;
    YA_CUDA(c, c->d_tb.reserve((size_t)tbUnits * 16 + 64));
    YA_CUDA(c, c->d_rows.reserve((size_t)rowsInts * 4 + 64));
    YA_CUDA(c, c->d_ops_raw.reserve((size_t)opsSlots * sizeof(ya_op) + 64));
    DpConst K{P.GOCost, P.GECost, P.RCost, P.MScore, P.XCutoff, P.maxIntron, P.maxGap, P.bandWidth};
    YA_CUDA(c, cudaEventRecord(c->ev[0], st));
    for (int k = 0; k < kNumWaveCfgs; k++) {
        if (count(k)) launch_wave_cfg(c, k, true, d_ids + start[k], count(k), K);
        if (count(kNumWaveCfgs + k)) launch_wave_cfg(c, k, false, d_ids + start[kNumWaveCfgs + k], count(kNumWaveCfgs + k), K);
    }
    const int packedBase = 2 * kNumWaveCfgs + 1;
    const int nP0 = count(packedBase), nP1 = count(packedBase + 1);
    if (nP0 || nP1) YA_CUDA(c, cudaEventRecord(c->ev[3], st));
    if (nP0) launch_packed_level(c, F.packedLevel, 0, d_ids + start[packedBase], nP0, K);
    if (nP1) launch_packed_level(c, F.packedLevel, 1, d_ids + start[packedBase + 1], nP1, K);
    if (nP0 || nP1) YA_CUDA(c, cudaEventRecord(c->ev[4], st));
    if (count(2 * kNumWaveCfgs)) {
        const int nt = count(2 * kNumWaveCfgs);
        dp_thread_kernel<<<(nt + 63) / 64, 64, 0, st>>>(c->d_jobs.as<DevJob>(), d_ids + start[2 * kNumWaveCfgs], nt, c->d_jobout.as<DevJobOut>(),
            c->d_tb.as<uint16_t>(), c->d_rows.as<int>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), K);
        c->ctr.launches++;
    }
    YA_CUDA(c, cudaEventRecord(c->ev[1], st));
    traceback_warp_kernel<<<(n_live + 3) / 4, 128, 0, st>>>(c->d_jobs.as<DevJob>(), d_ids, (int)n_live, c->d_jobout.as<DevJobOut>(),
        c->d_tb.as<uint16_t>(), c->d_ops_raw.as<ya_op>(), c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(),
        P.bandWidth, c->d_res.as<ya_dp_result>(), nullptr, d_acct);
    c->ctr.launches++;
    YA_CUDA(c, cudaEventRecord(c->ev[2], st));
    YA_CUDA(c, cudaGetLastError());
    c->dpr_ran = true;
    c->dpr_bulk_packed = (size_t)(nP0 + nP1) >= 4096;
    c->dpr_packed_launches = (nP0 ? 1 : 0) + (nP1 ? 1 : 0);
    return YA_OK;
}

// ------------------------------------------------------------------------------------------
// Perfect extension (AlignExtFrag.cpp:30-48)
// ------------------------------------------------------------------------------------------
__global__ void perfect_kernel(const ya_dp_job *__restrict__ jobs, const uint64_t *__restrict__ read_off, int n,
                               const uint8_t *__restrict__ bases, const uint8_t *__restrict__ fwd,
                               const uint8_t *__restrict__ rev, uint16_t *__restrict__ count)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const ya_dp_job j = jobs[t];
    const uint8_t *codes = (j.strand ? rev : fwd) + read_off[j.read];
    const int dir = (j.kind == YA_DP_EXT_BWD) ? -1 : 1;
    int k = 0;
    while (k < (int)j.qLen && codes[(int)j.qOff + dir * k] == nib(bases, j.rOff + (uint32_t)(dir * k))) k++;
    count[t] = (uint16_t)k;
}

extern "C" int ya_perfect_ext(ya_ctx *c, const ya_dp_job *jobs, int n, uint16_t *count)
{
    if (!c || n < 0 || (n && (!jobs || !count))) return YA_E_ARG;
    if (n == 0) return YA_OK;
    if (c->n_reads == 0) return ya_fail(c, YA_E_STATE, "ya_perfect_ext: no read batch uploaded");
    YA_CUDA(c, cudaSetDevice(c->device));
    // the walk must stay inside the read and inside the reference: clamp its length here, as the callers of
    // extendFragment*ToStopPerfectly do with min(query room, reference room) (AlignExtFrag.cpp:76-100)
    YA_CUDA(c, c->h_jobs.reserve((size_t)n * sizeof(ya_dp_job)));
    ya_dp_job *hj = c->h_jobs.as<ya_dp_job>();
    for (int i = 0; i < n; i++) {
        ya_dp_job j = jobs[i];
        if (j.read >= (uint32_t)c->n_reads || j.strand > 1 || (j.kind != YA_DP_EXT_FWD && j.kind != YA_DP_EXT_BWD))
            return ya_fail(c, YA_E_ARG, "ya_perfect_ext: bad job");
        const int L = (int)(c->h_read_off[j.read + 1] - c->h_read_off[j.read]);
        if ((int)j.qOff >= L || j.rOff > c->maxROff) return ya_fail(c, YA_E_ARG, "ya_perfect_ext: job starts outside the read or the reference");
        uint32_t room;
        if (j.kind == YA_DP_EXT_BWD) room = std::min<uint32_t>((uint32_t)j.qOff + 1u, j.rOff + 1u);
        else room = std::min<uint32_t>((uint32_t)(L - (int)j.qOff), c->maxROff - j.rOff + 1u);
        if ((uint32_t)j.qLen > room) j.qLen = (uint16_t)room;
        hj[i] = j;
    }
    AllocScope allocScope(c->stream);
    YA_CUDA(c, c->d_jobs.reserve((size_t)n * sizeof(ya_dp_job)));
    YA_CUDA(c, c->d_res.reserve((size_t)n * 2 + 64));
    YA_CUDA(c, cudaMemcpyAsync(c->d_jobs.p, hj, (size_t)n * sizeof(ya_dp_job), cudaMemcpyHostToDevice, c->stream));
    perfect_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->d_jobs.as<ya_dp_job>(), c->d_read_off.as<uint64_t>(), n,
        c->d_bases, c->d_codes_fwd.as<uint8_t>(), c->d_codes_rev.as<uint8_t>(), c->d_res.as<uint16_t>());
    c->ctr.launches++;
    YA_CUDA(c, cudaMemcpyAsync(count, c->d_res.p, (size_t)n * 2, cudaMemcpyDeviceToHost, c->stream));
    YA_CUDA(c, ya_stream_wait(c->stream));
    YA_CUDA(c, cudaGetLastError());
    return YA_OK;
}
