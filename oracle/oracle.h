/* oracle/oracle.h -- CPU restatement ("oracle") of the yaha 0.1.83 alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load liboracle.so; the product (yaha_b200/) never does.
 *
 * Parity status: PINNED.  Every function here is checked against outputs of the unmodified
 * reference compiled into oracle/_ref/ (see oracle/Makefile, oracle/dump_shim.c,
 * tests/test_oracle_vs_ref.py and the committed fixtures under tests/golden/).
 */
#ifndef YAHA_ORACLE_H
#define YAHA_ORACLE_H
#include "../include/yaha_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 4-bit code of reference base `off` (ref: getFrom4Code, Math.c:180-188). */
int orc_base(const uint8_t *bases, uint32_t off);

/* Forward codes from characters and reverse-complement codes (ref: Math.c:141-156,
 * Query.c:161-168). */
void orc_encode(const char *chars, int n, uint8_t *fwd, uint8_t *rev);

/* One DP job with the semantics of the findAGS* wrapper named by job->kind
 * (ref: SW.cpp:462-547 wrappers, SW.cpp:798-1208 DP + traceback).
 * `codes` is the strand's code buffer (forward or revcomp) of the job's read.
 * Returns the wrapper's return value; ops are written in genome order.
 * *cells gets the number of inner-loop bodies (SW.cpp:1007-1084) executed. */
int orc_dp(const ya_params *p, const uint8_t *bases, uint32_t maxROff, const uint8_t *codes,
           int kind, uint32_t rOff, int rLen, int qOff, int qLen,
           int *addedQLen, int *addedRLen, ya_op *ops, int ops_cap, int *n_ops, int64_t *cells);

/* Perfect extension count (ref: AlignExtFrag.cpp:30-48). dir: +1 forward, -1 backward. */
int orc_perfect(const uint8_t *bases, const uint8_t *codes, uint32_t rOff, int qOff, int len, int dir);

/* Stage 1 (ref: Query.c:365-412).  Fills sOffset[i], count[i] for i in [0, L-K]; count 0 for
 * dropped k-mers.  Returns totalCount. */
uint32_t orc_seed_lookup(const ya_params *p, const uint32_t *so, const uint8_t *codes, int L,
                         uint32_t *sOffset, uint32_t *count);

/* Stage 2a (ref: findFragmentsSort QueryMatch.c:52-121, heap QueryHeap.inl).  Returns fragCount
 * (or -1 if frag_cap is too small).  n_roa bounds the over-read quirk of QueryMatch.c:62-67. */
int orc_find_frags(const ya_params *p, const uint32_t *roa, size_t n_roa,
                   const uint32_t *sOffset, const uint32_t *count, int matchCount,
                   ya_frag *frags, int frag_cap);

/* Stage 2b (ref: QueryMatch.c:146-158, 224-303): region ordinal per fragment and the keep
 * flag (0 = singleton region with refLen < minMatch).  Returns number of regions. */
int orc_regions(const ya_params *p, const ya_frag *frags, int n, uint32_t *region, uint8_t *keep);

#ifdef __cplusplus
}
#endif
#endif
