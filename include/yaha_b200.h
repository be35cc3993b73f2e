/* yaha_b200.h -- C ABI of libyaha_b200.so: the B200-native replacement for the alignment
 * hot path of yaha 0.1.83 (seed lookup -> diagonal sort / fragments / regions -> banded
 * affine-gap DP with X-drop).  Plain C, POD structs, pointers and sizes only.
 *
 * Every entry point names the reference interface it replaces ("ref:" = file:line under
 * /root/reference/src).  All buffers passed in are HOST memory owned by the caller; the
 * library owns every device allocation.  A ya_ctx is bound to one GPU and must be used from
 * one host thread at a time (the reference's per-thread QueryState_t plays the same role,
 * Math.h:587-666).  There is no CPU fallback: if no usable sm_100 device is present ya_open
 * fails and every other call returns YA_E_CUDA.
 */
#ifndef YAHA_B200_H
#define YAHA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (ref: errors are fatal exit(1), FileHelpers.c:38-42; here they return) ---- */
#define YA_OK          0
#define YA_E_ARG       1   /* bad argument / unsupported parameter combination            */
#define YA_E_CUDA      2   /* CUDA runtime error, text in ya_last_error()                 */
#define YA_E_CAPACITY  3   /* a caller-provided output buffer is too small; *_needed set  */
#define YA_E_STATE     4   /* call order violated (e.g. no reads uploaded), or -- ya_align_batch -- the batch does not
                              fit one device pass (caller: smaller batches, or the call-by-call path)              */
#define YA_E_INTERNAL  5   /* a consistency check on the device failed (a bug: never retried on another path)      */

/* ---- scoring / seeding parameters: POD copy of the AlignmentArgs_t fields the hot path
 *      reads (ref: Math.h:281-304, defaults AlignArgs.c:48-87, derived :108-169) ---- */
typedef struct ya_params {
    int32_t wordLen;       /* K, from the index header (Query.c:603)                      */
    int32_t maxHits;       /* min(-H or 650, index header) (Query.c:604-610)              */
    int32_t bandWidth;     /* -BW                                                         */
    int32_t maxGap;        /* -G  : cap on insert run length inside the DP (SW.cpp:1050)  */
    int32_t maxIntron;     /* = maxGap unless set: cap on delete run length (SW.cpp:1032) */
    int32_t minMatch;      /* -M  : singleton-region filter (QueryMatch.c:284)            */
    int32_t GOCost, GECost, RCost, MScore;   /* -GOC -GEC -RC -MS                         */
    int32_t XCutoff;       /* -X                                                          */
    int32_t minExtLength;  /* derived, AlignArgs.c:141-149 (host-side dispatch only)      */
} ya_params;

/* ---- Fragment: bit-compatible with the reference's Fragment_t (Math.h:448-455) ---- */
typedef struct ya_frag {
    uint32_t startRefOff;
    uint16_t startQueryOff;
    uint16_t endQueryOff;
    uint16_t hitCount;     /* never written by the reference (QueryMatch.c:52-121); 0 here */
    uint16_t refLen;
} ya_frag;

/* ---- a batch of reads: forward strand 4-bit codes, one byte per base (the reference's
 *      QS->forwardCodeBuf, Query.c:161-163).  The reverse-complement strand
 *      (Query.c:164-167) is derived on the device. ---- */
typedef struct ya_read_batch {
    int32_t         n_reads;
    const uint8_t  *codes;      /* concatenated codes, offsets[n_reads] bytes              */
    const uint64_t *offsets;    /* n_reads+1 entries, offsets[0] == 0                      */
} ya_read_batch;

/* ---- stage 1+2 output.  For each (read,strand) the diag-sorted fragment array of
 *      findFragmentsSort (QueryMatch.c:52) after the region scan of processFragmentsGapped
 *      (QueryMatch.c:224-303): fragments of singleton regions with refLen < minMatch are
 *      dropped on the device (they can never reach a clump, QueryMatch.c:281-290); every
 *      other fragment is returned, in the reference's order, with its region id. ---- */
typedef struct ya_strand_frags {
    uint32_t first;        /* index of this strand's first surviving fragment in frags[]  */
    uint32_t n_frags;      /* surviving fragments                                          */
    uint32_t n_frags_all;  /* fragCount the reference would have returned                  */
    uint32_t total_hits;   /* totalCount of Query.c:365-412 (sum of kept k-mer counts)    */
} ya_strand_frags;

typedef struct ya_frag_batch {
    /* capacities, set by the caller */
    size_t           frags_cap;
    /* outputs */
    ya_strand_frags *strands;     /* [2*n_reads]: index 2*r + strand (0 fwd, 1 revcomp)   */
    ya_frag         *frags;       /* [frags_cap]                                           */
    uint32_t        *region;      /* [frags_cap] region ordinal within its strand; a       */
                                  /* region with one member is a singleton (-> clump),     */
                                  /* otherwise it goes to the host graph (GraphPath.cpp)   */
    size_t           n_frags;     /* total surviving fragments written                     */
    size_t           frags_needed;/* set when YA_E_CAPACITY is returned                    */
} ya_frag_batch;

/* ---- stage 3: DP jobs.  Semantics of each kind are those of the reference wrapper named,
 *      including reference-end clamping (SW.cpp:496-516) and the "<= 0 means no extension,
 *      no ops" rule (SW.cpp:525,1102). ---- */
#define YA_DP_FULL     0   /* findAGSAlignment         Math.h:401  SW.cpp:462              */
#define YA_DP_BANDED   1   /* findAGSAlignmentBanded   Math.h:402  SW.cpp:470              */
#define YA_DP_EXT_FWD  2   /* findAGSForwardExtension  Math.h:403  SW.cpp:543              */
#define YA_DP_EXT_BWD  3   /* findAGSBackwardExtension Math.h:407  SW.cpp:537              */

typedef struct ya_dp_job {
    uint32_t rOff;         /* reference offset argument of the wrapper                     */
    uint32_t read;         /* read index in the uploaded batch                             */
    uint16_t rLen;         /* FULL/BANDED only                                             */
    uint16_t qOff;
    uint16_t qLen;
    uint8_t  kind;         /* YA_DP_*                                                      */
    uint8_t  strand;       /* 0 = forward codes, 1 = reverse-complement codes              */
} ya_dp_job;

typedef struct ya_dp_result {
    int32_t  score;        /* return value of the wrapper                                  */
    uint16_t addedQLen;    /* extensions only (SW.cpp:1109)                                */
    uint16_t addedRLen;    /* extensions only (SW.cpp:1110)                                */
    uint32_t ops_off;      /* first op of this job in ops[]                                */
    uint32_t ops_n;        /* number of run-length ops (0 when score <= 0 for extensions) */
} ya_dp_result;

/* run-length edit op in genome order (ref: EditOp_t Math.h:371-378 without the links) */
typedef struct ya_op {
    uint16_t length;
    uint8_t  opcode;       /* 'M' 'R' 'I' 'D' (Math.h:354-360)                             */
    uint8_t  pad;
} ya_op;

typedef struct ya_counters {
    uint64_t probes;       /* k-mer probes issued (stage 1)                                */
    uint64_t hits;         /* seed hits expanded (stage 2a)                                */
    uint64_t frags_all;    /* fragments formed                                             */
    uint64_t frags_out;    /* fragments returned to the host                               */
    uint64_t dp_jobs;
    uint64_t dp_cells;     /* cells as the reference counts them (SW.cpp:1007-1084 bodies) */
    double   ms_seed;      /* device time of stage 1+2 kernels (CUDA events)               */
    double   ms_dp;        /* device time of stage 3 fill kernels                          */
    double   ms_traceback; /* device time of stage 3 traceback kernels                     */
    uint64_t launches;     /* kernels launched by this library                             */
    uint64_t ext_cells;    /* cells of the bulk (>= 4096 jobs) dp_ext_packed_kernel launches */
    double   ms_ext;       /* device time of dp_ext_packed_kernel launches (CUDA events)      */
    uint64_t ext_launches; /* number of dp_ext_packed_kernel launches in ms_ext               */
    double   ms_lookup;    /* device time of seed_count_kernel (the SO gathers) alone         */
    double   ms_finish;    /* device time of the assemble / finish / format kernels (ya_align_batch) */
    uint64_t reads_finished, reads_handed_back;   /* ya_align_batch: reads done on the device / returned with status 1 */
    uint64_t text_bytes;   /* SAM text produced on the device                                 */
} ya_counters;

typedef struct ya_ctx ya_ctx;

/* Create a context on CUDA device `device` and make the index + genome resident in HBM.
 *   so  : the index's starting-offset table, 4^K+1 uint32   (ref: AAs->startingOffs, Query.c:625)
 *   roa : the index's reference-offset array, n_roa uint32  (ref: AAs->ROAPtr, Query.c:626)
 *   bases: .nib2 base area, 4 bits per base, high nibble first (ref: AAs->basePtr, Query.c:572)
 *   maxROff: baseSequencesMaxROff (BaseSeq.c:121-125)
 * Returns NULL on failure; ya_last_error(NULL) describes it. */
ya_ctx *ya_open(int device, const ya_params *params,
                const uint32_t *so, size_t n_so, const uint32_t *roa, size_t n_roa,
                const uint8_t *bases, size_t n_base_bytes, uint32_t maxROff);

/* Same, but clone the device-resident index of `src` (another GPU) by peer copy over
 * NVLink instead of re-uploading from the host. */
ya_ctx *ya_open_peer(int device, const ya_ctx *src);
/* 1 if the replica of this context was copied with peer access enabled in both directions (a direct HBM-to-HBM copy over
 * NVLink / NVSwitch), 0 if the driver had to stage it through the host or the context was not made by ya_open_peer. */
int ya_peer_direct(const ya_ctx *);

/* Same as ya_open, but BUILD the index on the device from the .nib2 base area instead of loading
 * an index file (replaces indexFile, Index.c:49-331, for every -L / -S / -H).  seq_start/seq_len are the
 * global base offset and length of each sequence (BaseSeq.c:115-119); index_max_hits and skip_dist are the
 * -H and -S of `yaha -g`.  The result is bit-identical to the reference's file contents, including the
 * down-sampling of k-mers with more than index_max_hits occurrences (Index.c:271-315: Floyd sampling from one
 * xorshift stream, Math.c:274-343 -- the stream is sequential, so its draws are made on the host for the few
 * over-full lists the device finds); ya_index_download() returns it for writing in the reference's format. */
ya_ctx *ya_open_build(int device, const ya_params *params, const uint8_t *bases, size_t n_base_bytes,
                      const uint32_t *seq_start, const uint32_t *seq_len, int n_seq, uint32_t index_max_hits,
                      uint32_t skip_dist);
int     ya_index_sizes(const ya_ctx *, size_t *n_so, size_t *n_roa);
int     ya_index_download(ya_ctx *, uint32_t *so, uint32_t *roa);

/* A second context on the SAME device that borrows src's resident index (no copy) but has its own
 * stream, read batch and scratch, so that two host pipelines can overlap their device and host
 * phases (one QueryState_t per thread in the reference, Query.c:660-667).  src must outlive it. */
ya_ctx *ya_open_shared(const ya_ctx *src);

void        ya_close(ya_ctx *);
const char *ya_last_error(const ya_ctx *);

/* Change scoring parameters without reloading the index (wordLen must not change). */
int ya_set_params(ya_ctx *, const ya_params *);

/* Run subsequent work on a caller-provided cudaStream_t (e.g. torch's current stream) so
 * that the caller's CUDA events bracket the kernels.  NULL restores the library's stream. */
int ya_set_stream(ya_ctx *, void *cuda_stream);

/* Make the context's device the calling thread's current CUDA device.  A host thread that allocates page-locked memory
 * (ya_host_alloc) before it has made any call with a context would otherwise do so against device 0 and create a CUDA
 * context there -- in every process of a multi-GPU job. */
int ya_bind_thread(const ya_ctx *);

/* Urgency of this context's work among the contexts that share its device (pipelines of one host program): level 0 (default)
 * = most urgent, up to 3.  Maps to CUDA stream priorities; call between batches.  The host gives the batches in flight
 * descending urgency in input order so that they finish one after the other and the ordered writer overlaps with compute. */
int ya_set_priority(ya_ctx *, int level);

/* Page-locked host memory for the buffers handed to the calls below (optional: any host pointer
 * works, page-locked ones are copied by DMA without an intermediate staging copy).  NULL on failure. */
void *ya_host_alloc(size_t bytes);
void  ya_host_free(void *p);

/* Make a batch of reads resident on the device (replaces the per-read buffers filled by
 * readNextQuery, Query.c:161-168).  Stays valid until the next ya_reads_upload. */
int ya_reads_upload(ya_ctx *, const ya_read_batch *);

/* Stage 1 + 2 for the uploaded batch.  Replaces, per (read,strand): the seed-lookup loop
 * Query.c:365-412, findFragmentsSort (Math.h:554, QueryMatch.c:52-121) and the region scan /
 * singleton filter of processFragmentsGapped (Math.h:555, QueryMatch.c:224-303). */
int ya_seed_frags(ya_ctx *, ya_frag_batch *out);

/* ---- Next row N1: fragments -> clumps of seed fragments on the device.
 * Replaces, for the batch ya_seed_frags has just processed, the region loop of processFragmentsGapped
 * (QueryMatch.c:224-303), processFragmentRangeUsingGraph / buildBestClumpFromFragmentRange
 * (GraphPath.cpp:161-292), eliminateFragments (QueryMatch.c:170-215) and addFragment / insertFragment /
 * cleanUpClump (AlignHelpers.c:48-193).  One thread per strand runs yaha_b200/csrc/form_clumps.h -- the
 * same source the host program compiles for its own formClumps.
 * Output, per strand s (index 2*r + strand as in ya_frag_batch): clumps
 * clumps[clump_first[s] .. clump_first[s] + clump_count[s]) in the reference's creation order; a clump's
 * fragments (query order, after overlap chops and clean-up) are path[rec.first .. rec.first + rec.n).
 * Capacities: clumps and path hold at most as many entries as ya_seed_frags returned fragments. ---- */
typedef struct ya_clump_rec {
    uint32_t first;        /* first fragment of the clump in path[]                              */
    uint16_t n;            /* fragments in the clump                                              */
    uint16_t matchedBases; /* Clump_t.matchedBases after addFragment (16-bit, as in the reference) */
} ya_clump_rec;

typedef struct ya_clump_batch {
    /* parameters that ya_params does not carry (AlignmentArgs_t.maxDesert, .minNonOverlap)      */
    int32_t       maxDesert, minNonOverlap;
    /* capacity, set by the caller: entries in clumps[] and in path[] (>= ya_frag_batch.n_frags) */
    size_t        cap;
    /* outputs */
    uint32_t     *clump_first;   /* [2*n_reads]                                                  */
    uint32_t     *clump_count;   /* [2*n_reads]                                                  */
    ya_clump_rec *clumps;        /* [cap]                                                        */
    ya_frag      *path;          /* [cap]                                                        */
    size_t        n_clumps, n_path;
} ya_clump_batch;

int ya_form_clumps(ya_ctx *, ya_clump_batch *out);

/* ---- First phase of alignClump on the device, for the clumps ya_form_clumps has just made: perfect extensions
 * between neighbouring seed fragments (AlignHelpers.c:226-237, AlignExtFrag.cpp:30-48), the dispatch of
 * makeAndAlignSFragmentToFillGap for every gap (closed form: pure D / pure I / 1x1 R; else a banded or full DP job,
 * AlignExtFrag.cpp:164-234) and the plan of the two end extensions (AlignExtFrag.cpp:64-107).  Same source as the
 * host program's own phase 1 (yaha_b200/csrc/prepare_clumps.h).  Records are indexed like ya_clump_batch:
 * prep[k] belongs to clumps[k]; its gaps are gaps[prep[k].gap_first .. + n_gaps); path[] holds the clumps'
 * fragments after the perfect extensions (same positions as ya_clump_batch.path).  jobs[] is ready for
 * ya_sw_batch; gap and extension records carry the index of their job (0xFFFFFFFF: none). ---- */
typedef struct ya_gap_rec {
    uint32_t job;          /* index into jobs[], or 0xFFFFFFFF for a closed-form gap                  */
    int32_t  score;        /* closed form: score of the piece                                          */
    uint16_t after;        /* the gap follows fragment `after` of its clump                            */
    uint16_t len;          /* closed form: run length                                                  */
    uint8_t  code;         /* closed form: 'D', 'I' or 'R'                                             */
    uint8_t  pad; uint16_t pad2;
} ya_gap_rec;

typedef struct ya_prep_rec {
    uint32_t gap_first;
    uint32_t jobB, jobF;   /* backward / forward extension job, 0xFFFFFFFF when shorter than minExtLength */
    uint16_t n_gaps;
    uint16_t backLen, forwLen;   /* lengths left for the DP after the perfect pre-extension             */
    uint16_t pad;
} ya_prep_rec;

typedef struct ya_prep_batch {
    size_t       cap;        /* entries in prep[], gaps[], path[] (>= ya_frag_batch.n_frags)            */
    size_t       jobs_cap;   /* entries in jobs[] (3 * cap always suffices)                             */
    ya_prep_rec *prep;
    ya_gap_rec  *gaps;
    ya_frag     *path;
    ya_dp_job   *jobs;
    size_t       n_jobs;
} ya_prep_batch;

int ya_prepare_clumps(ya_ctx *, ya_prep_batch *out);

/* ---- Rows N2 + N4: the whole per-read path for a batch in ONE call.  Replaces, per read, the body of
 * the processQueries loop (Query.c:306-497) and printClumps (QueryMatch.c:333-344, AlignOutput.c:115-289): the reads go up
 * as they stand in the query file (characters, ids); the device encodes them (Query.c:161-168), runs stages 1-3 with the
 * jobs of the first DP round born, laid out and answered on the device, splices and scores every clump
 * (AlignHelpers.c:251-366), runs Optimal Query Coverage / filter by similarity (GraphPath.cpp:294-1086) and writes the SAM
 * records (AlignOutput.c:115-289) of every read it can finish, in read order, into one text buffer.
 * A read is handed back (status 1, no text) when the reference's control flow for it leaves the straight path: a clump
 * splitClump has to look at (AlignHelpers.c:374-579), a strand too crowded for the device's fragment graph, more scored
 * clumps than the device's graph holds, -OQC N with several clumps.  The caller runs those reads through the calls above.
 * ya_set_output must have been called.  On YA_E_CAPACITY (text_cap too small) text_needed is set, status / text_off are
 * valid and the text stays on the device for ya_align_fetch_text. ---- */
typedef struct ya_out_params {       /* the AlignmentArgs_t fields the tail of the per-read path reads (Math.h:257-334) */
    int32_t maxDesert, minNonOverlap;                        /* fragment graph (GraphPath.cpp:161-292)          */
    int32_t minRawScore;  float minIdentity;                 /* scoreClump thresholds (AlignHelpers.c:343-361)   */
    int32_t OQC, FBS, OQCMinNonOverlap, BPCost, maxBPLog;    /* -OQC -FBS -MNO -BP -MGDP                        */
    float   FBS_PSLength, FBS_PSScore;                       /* -PRL -PSS                                       */
    int32_t hardClip, fastq;                                 /* -osh / -oss; FASTQ input (QUAL column)           */
} ya_out_params;

/* names / starts / lengths of the reference sequences (BaseSequence_t, Math.h:218-225; the @SQ lines and the RNAME column) */
int ya_set_output(ya_ctx *, const ya_out_params *, int n_seq, const char *const *seq_names,
                  const uint32_t *seq_start, const uint32_t *seq_len);

typedef struct ya_text_batch {
    int32_t         n_reads;
    const char     *chars;      /* the reads' characters as in the file, concatenated                      */
    const uint64_t *offsets;    /* n_reads+1 entries, offsets[0] == 0                                      */
    const char     *quals;      /* FASTQ: quality characters, same offsets; NULL for FASTA                 */
    const char     *ids;        /* ids (<= 200 characters, blanks replaced: Query.c:111-135), concatenated */
    const uint32_t *id_off;     /* n_reads+1 entries                                                       */
    /* outputs (caller's buffers) */
    char           *text;       /* SAM records of the reads finished on the device, in read order          */
    size_t          text_cap;
    uint64_t       *text_off;   /* [n_reads+1]: read r's records are text[text_off[r] .. text_off[r+1])     */
    uint8_t        *status;     /* [n_reads] 0: finished here (possibly without a record), 1: handed back   */
    size_t          text_len, text_needed;
    int32_t         n_handed_back;
} ya_text_batch;

int ya_align_batch(ya_ctx *, ya_text_batch *);
int ya_align_fetch_text(ya_ctx *, char *text, size_t text_cap);

/* Stage 3 for n independent jobs against the uploaded batch.  Replaces findAGSAlignment,
 * findAGSAlignmentBanded, findAGSForwardExtension, findAGSBackwardExtension
 * (Math.h:401-408) = findAffineGapScore<...> (SW.cpp:798-1208) + decompressRef
 * (SW.cpp:444-456).  res[i] answers jobs[i].  On YA_E_CAPACITY *ops_needed is set, every
 * res[i] is already valid and the edit operations stay on the device until the next
 * ya_sw_batch: collect them with ya_sw_fetch_ops (no DP is repeated), or repeat the call. */
int ya_sw_batch(ya_ctx *, const ya_dp_job *jobs, int n, ya_dp_result *res,
                ya_op *ops, size_t ops_cap, size_t *ops_needed);
/* Copy the edit operations of the last ya_sw_batch that returned YA_E_CAPACITY (ops_needed
 * entries).  YA_E_CAPACITY again if ops_cap is still too small, YA_E_STATE if nothing is pending. */
int ya_sw_fetch_ops(ya_ctx *, ya_op *ops, size_t ops_cap);

/* Perfect (exact-match) extension lengths (ref: extendFragment{Forward,Backward}
 * ToStopPerfectly, AlignExtFrag.cpp:30-48): for job i, count equal codes starting at
 * (qOff,rOff) walking forward (kind EXT_FWD) or backward (EXT_BWD), at most qLen. */
int ya_perfect_ext(ya_ctx *, const ya_dp_job *jobs, int n, uint16_t *count);

/* Read-and-reset the counters. */
int ya_get_counters(ya_ctx *, ya_counters *);

/* The [start, end) spans, in ms on a time axis common to all contexts of the device, of the bulk (>= 4096 jobs)
 * dp_ext_packed_kernel launches since the last call (pairs of floats).  The pipelines of a device overlap their launches; the
 * union of the spans is the time the kernel class occupied the device.  YA_E_CAPACITY: *n_pairs says how many there are. */
int ya_get_ext_intervals(ya_ctx *, float *start_end_ms, int cap_pairs, int *n_pairs);

/* Device-side self measurements used by bench.py for roofline denominators. */
/* Sustained INT32 issue rate of this GPU in 1e9 lane-operations per second, all SMs busy:
 * *giops_add from a pure dependent-free IADD3 stream, *giops_mix from the add / compare /
 * select / min-max mix of the DP cell (the roofline denominator of the banded-SW kernel). */
int ya_measure_int32_peak(ya_ctx *, double *giops_add, double *giops_mix);

/* Independent random 4-byte gathers per second over the resident starting-offset table (all SMs, 8 in
 * flight per thread): the HBM sector-miss rate that bounds the seed lookup (Query.c:391). */
int ya_measure_gather_peak(ya_ctx *, double *gather_per_s);

#ifdef __cplusplus
}
#endif
#endif /* YAHA_B200_H */
