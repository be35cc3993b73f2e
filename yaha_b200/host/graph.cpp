// graph.cpp -- fragments -> clumps.  The algorithm itself (region loop, fragment graph, overlap chops, clean-up,
// elimination: QueryMatch.c:224-303, GraphPath.cpp:57-292, AlignHelpers.c:48-193) is stated once, in
// yaha_b200/csrc/form_clumps.h, and compiled here for the host and in clumps.cu for the device.  This file
// turns its output -- or the device's, when the pipeline asked for it (ya_form_clumps) -- into Clump objects.
#include <algorithm>
#include <string.h>
#include "host.hpp"
#include "../csrc/form_clumps.h"

namespace yh {

void disposeClump(Clump *c) { delete c; }

static fc_params clumpParams(const Args &A)
{
    fc_params P;
    P.wordLen = A.wordLen; P.maxGap = A.maxGap; P.maxDesert = A.maxDesert; P.minMatch = A.minMatch;
    P.minNonOverlap = A.minNonOverlap; P.bandWidth = A.bandWidth; P.GOCost = A.GOCost; P.GECost = A.GECost; P.MScore = A.MScore;
    return P;
}

// Clump objects from clump records: the clumps of a strand in creation order (addClump, QueryState.c:156-161)
static void makeClumps(const Env &E, ReadCtx &rc, bool rev, const ya_clump_rec *recs, int nClumps, const Frag *path,
                       const ya_prep_rec *prep = nullptr, const ya_gap_rec *gapBase = nullptr)
{
    const uint8_t *bases = E.G->bases;
    for (int k = 0; k < nClumps; k++) {
        Clump *c = new Clump();
        const Frag *p = path + recs[k].first;
        c->path.assign(p, p + recs[k].n);
        c->matchedBases = recs[k].matchedBases;
        c->set(kReversed, rev);
        if (prep) { c->prep = prep + k; c->gapBase = gapBase; }
        // the alignment phase that follows compares bases just outside both ends of every piece (perfect extensions,
        // AlignExtFrag.cpp:30-48): ask for those genome lines now, they are cache misses in a 50 MB .. 1.5 GB array
        // (with phase 1 done on the device only the clump's two outer ends are still compared here)
        for (int q = 0; q < (int)recs[k].n; q++) {
            if (prep && q != 0 && q != (int)recs[k].n - 1) continue;
            if (!prep || q == 0) __builtin_prefetch(bases + ((p[q].startRefOff - 1u) >> 1));
            if (!prep || q == (int)recs[k].n - 1) __builtin_prefetch(bases + ((p[q].startRefOff + p[q].refLen) >> 1));
        }
        rc.clumps.push_back(c);
    }
}

void formClumps(const Env &E, ReadCtx &rc, bool rev)
{
    const int n = rc.nFrags[rev];
    if (rc.devClumps[rev]) {                                    // formed on the device (ya_form_clumps)
        makeClumps(E, rc, rev, rc.devClumps[rev], rc.nDevClumps[rev], rc.devPath[rev], rc.devPrep[rev], rc.devGaps);
        return;
    }
    if (n == 0) return;
    static thread_local std::vector<fc_node> nodes;
    static thread_local std::vector<uint8_t> used;
    static thread_local std::vector<Frag> tmp, path;
    static thread_local std::vector<ya_clump_rec> recs;
    if ((int)nodes.size() < n) { nodes.resize((size_t)n); tmp.resize((size_t)n); path.resize((size_t)n); recs.resize((size_t)n); used.resize((size_t)2 * n); }
    const fc_params P = clumpParams(*E.A);
    const int nClumps = fc_form_clumps(&P, rc.frags[rev], rc.region[rev], n, rc.read->len(), nodes.data(), used.data(), tmp.data(),
                                       path.data(), recs.data());
    makeClumps(E, rc, rev, recs.data(), nClumps, path.data());
}

}  // namespace yh
