/* oracle_clumps.c -- TEST INFRASTRUCTURE.  CPU build of yaha_b200/csrc/form_clumps.h (fragments of one strand ->
 * clumps of seed fragments: QueryMatch.c:224-303, GraphPath.cpp:161-292, AlignHelpers.c:48-193), the source the
 * device kernel (clumps.cu) and the host program (host/graph.cpp) compile as well.  It is pinned through the host
 * program: with this code in formClumps the host reproduces every golden SAM of the unmodified reference
 * (tests/test_host_mock.py); tests/test_gpu_parity.py compares the kernel's records with this build strand by strand. */
#include <stdlib.h>
#include <string.h>
#include "../yaha_b200/csrc/form_clumps.h"
#include "../yaha_b200/csrc/assemble_clumps.h"

int orc_form_clumps(int wordLen, int maxGap, int maxDesert, int minMatch, int minNonOverlap, int bandWidth,
                    int GOCost, int GECost, int MScore,
                    const ya_frag *frags, const uint32_t *region, int n, int readLen,
                    ya_frag *out_path, ya_clump_rec *out_clumps)
{
    fc_params P;
    P.wordLen = wordLen; P.maxGap = maxGap; P.maxDesert = maxDesert; P.minMatch = minMatch; P.minNonOverlap = minNonOverlap;
    P.bandWidth = bandWidth; P.GOCost = GOCost; P.GECost = GECost; P.MScore = MScore;
    if (n <= 0) return 0;
    ya_frag *work = malloc((size_t)n * sizeof(ya_frag)), *tmp = malloc((size_t)n * sizeof(ya_frag));
    fc_node *nodes = malloc((size_t)n * sizeof(fc_node));
    uint8_t *used = malloc(2 * (size_t)n);
    memcpy(work, frags, (size_t)n * sizeof(ya_frag));
    const int nc = fc_form_clumps(&P, work, region, n, readLen, nodes, used, tmp, out_path, out_clumps);
    free(work); free(tmp); free(nodes); free(used);
    return nc;
}

/* CPU build of yaha_b200/csrc/assemble_clumps.h (what follows a clump's first DP round: splice, both end extensions,
 * scoreClump's verdict).  tests/test_assemble_clumps.py compares it with a step-by-step restatement of the reference's
 * list operations (five merges with mergeEOLToFront / mergeEOLToBack, then the walk of AlignHelpers.c:302-366). */
int orc_assemble_clump(int GOCost, int GECost, int RCost, int MScore, int minExtLength, int minRawScore, uint32_t maxROff,
                       double minIdentity, const uint8_t *bases, const uint8_t *q, int readLen,
                       const ya_frag *p, int np, const ya_gap_rec *gaps, int ng, const ya_prep_rec *prep,
                       const ya_dp_result *res, const ya_op *rops, ya_op *out, uint32_t out_cap, ya_asm_rec *rec)
{
    ac_params P;
    P.GOCost = GOCost; P.GECost = GECost; P.RCost = RCost; P.MScore = MScore; P.minExtLength = minExtLength;
    P.minRawScore = minRawScore; P.maxROff = maxROff; P.minIdentity = minIdentity;
    if (ac_ops_bound(np, gaps, ng, prep, res, rops) > out_cap) return -2;
    return ac_assemble_clump(&P, bases, q, readLen, p, np, gaps, ng, prep, res, rops, out, rec);
}
