// args.cpp -- command line, defaults and derived parameters (host-kept CLI surface).
// Follows: makeAlignmentArgs AlignArgs.c:27-89, postProcessAlignmentArgs AlignArgs.c:108-169,
//          main() flag cascade and file-name derivation Main.c:187-565.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "host.hpp"

namespace yh {

void Args::postProcess(bool queryMode)
{
    if (maxIntron == -1) maxIntron = maxGap;
    if (minRawScore == -1) minRawScore = minMatch;
    if (OQCMinNonOverlap == -1) OQCMinNonOverlap = minMatch;
    if (OQCMinNonOverlap <= 0) { fprintf(stderr, "MNO parameter must be >=1.  MNO=1 will be used.\n"); OQCMinNonOverlap = 1; }
    if (minNonOverlap == -1) minNonOverlap = OQCMinNonOverlap;
    if (!affineGapScoring) { MScore = 1; RCost = GECost = 1; GOCost = 0; }
    int len = 1, score = 0, target = std::min(RCost, GOCost + GECost);
    while (score <= target) { score += MScore; len += 1; }
    minExtLength = (uint8_t)len;
    if (maxHits == -1) maxHits = queryMode ? 650 : 65525;
    else maxHits = std::min(maxHits, 65525);
    if (maxBPLog < 1) { fprintf(stderr, "MGDP parameter must be between 1 and 9 (inclusive). MGDP=1 will be used.\n"); maxBPLog = 1; }
    if (maxBPLog > 9) { fprintf(stderr, "MGDP parameter must be between 1 and 9 (inclusive). MGDP=9 will be used.\n"); maxBPLog = 9; }
}

ya_params Args::deviceParams() const
{
    ya_params p;
    p.wordLen = wordLen; p.maxHits = maxHits; p.bandWidth = bandWidth; p.maxGap = maxGap; p.maxIntron = maxIntron;
    p.minMatch = minMatch; p.GOCost = GOCost; p.GECost = GECost; p.RCost = RCost; p.MScore = MScore; p.XCutoff = XCutoff;
    p.minExtLength = minExtLength;
    return p;
}

static void usage()
{
    fprintf(stderr,
        "Usage (same flags as yaha 0.1.83):\n"
        "  index : yaha_b200 -g genome.(fa|nib2) [-L wordLen] [-S skipDist] [-H maxHits]\n"
        "  align : yaha_b200 -x index -q reads.(fa|fq) [-osh|-oss|-o8 out] [-t threads]\n"
        "          [-BW n] [-G n] [-H n] [-M n] [-MD n] [-P f] [-X n] [-AGS Y|N] [-GEC n] [-GOC n] [-MS n] [-RC n]\n"
        "          [-OQC Y|N] [-BP n] [-MGDP n] [-MNO n] [-FBS Y|N] [-PRL f] [-PSS f]\n"
        "  yaha_b200 only: [-gpus N] [-dev first-device] [-batch reads-in-flight] [-passes N] [-pipes P] [-replay]\n");
}

static bool parseBool(const char *s, const char *key)
{
    if (strlen(s) == 1) {
        if (strchr("YyTt", s[0])) return true;
        if (strchr("NnFf", s[0])) return false;
    }
    fprintf(stderr, "%s is not a valid value for parameter %s.\nUse one of 'YyTt' for Yes and 'NnFf' for No.\n\n.", s, key);
    usage(); exit(1);
}

static int parseInt(const char *s, const char *key)
{
    int v = atoi(s);
    if (v < 0) { fprintf(stderr, "%s is not a valid value for parameter %s.\nValue must be a positive integer.\n\n", s, key); usage(); exit(1); }
    return v;
}

static float parseFloat(const char *s, const char *key)
{
    float v = (float)atof(s);
    if (v <= 0.0 || v > 1.0) { fprintf(stderr, "%s is not a valid value for parameter %s.\nValue must be in the range 0<value<=1.0.\n\n", s, key); usage(); exit(1); }
    return v;
}

int parseArgs(int argc, char **argv, Args &a)
{
    if (argc <= 1) { usage(); return 1; }
    bool query = false, index = true;
    for (int x = 1; x < argc; x++) {
        const char *k = argv[x];
        auto val = [&]() -> const char * { if (x + 1 >= argc) { fprintf(stderr, "%s needs a value.\n", k); usage(); exit(1); } return argv[++x]; };
        if (!strcmp(k, "-h") || !strcmp(k, "-?") || !strcmp(k, "-xh")) { usage(); return 1; }
        else if (!strcmp(k, "-g")) { a.gfile = val(); a.haveG = true; }
        else if (!strcmp(k, "-q")) {
            const char *v = val();
            a.qfile = (!strcmp(v, "-") || !strcmp(v, "-stdin") || !strcmp(v, "stdin")) ? "stdout" : v;     // Main.c:173-178 (sic)
            query = true; index = false;
        }
        else if (!strcmp(k, "-o8")) { a.outputBlast8 = true; a.outputSAM = false; const char *v = val(); a.ofile = (!strcmp(v, "-stdout")) ? "stdout" : v; a.haveO = true; }
        else if (!strcmp(k, "-osh")) { a.outputBlast8 = false; a.outputSAM = true; a.hardClip = true; const char *v = val(); a.ofile = (!strcmp(v, "-stdout")) ? "stdout" : v; a.haveO = true; }
        else if (!strcmp(k, "-oss")) { a.outputBlast8 = false; a.outputSAM = true; a.hardClip = false; const char *v = val(); a.ofile = (!strcmp(v, "-stdout")) ? "stdout" : v; a.haveO = true; }
        else if (!strcmp(k, "-t")) a.numThreads = parseInt(val(), "-t");
        else if (!strcmp(k, "-v")) a.verbose = true;
        else if (!strcmp(k, "-x")) { a.xfile = val(); a.haveX = true; query = true; index = false; }
        else if (!strcmp(k, "-H")) a.maxHits = parseInt(val(), "-H");
        else if (!strcmp(k, "-L")) a.wordLen = parseInt(val(), "-L");
        else if (!strcmp(k, "-S")) a.skipDist = parseInt(val(), "-S");
        else if (!strcmp(k, "-BW")) a.bandWidth = parseInt(val(), "-BW");
        else if (!strcmp(k, "-G")) a.maxGap = parseInt(val(), "-G");
        else if (!strcmp(k, "-M")) a.minMatch = parseInt(val(), "-M");
        else if (!strcmp(k, "-MD")) a.maxDesert = parseInt(val(), "-MD");
        else if (!strcmp(k, "-P")) a.minIdentity = parseFloat(val(), "-P");
        else if (!strcmp(k, "-X")) a.XCutoff = parseInt(val(), "-X");
        else if (!strcmp(k, "-AGS")) a.affineGapScoring = parseBool(val(), "-AGS");
        else if (!strcmp(k, "-GEC")) a.GECost = parseInt(val(), "-GEC");
        else if (!strcmp(k, "-GOC")) a.GOCost = parseInt(val(), "-GOC");
        else if (!strcmp(k, "-MS")) a.MScore = parseInt(val(), "-MS");
        else if (!strcmp(k, "-RC")) a.RCost = parseInt(val(), "-RC");
        else if (!strcmp(k, "-OQC")) a.OQC = parseBool(val(), "-OQC");
        else if (!strcmp(k, "-BP")) a.BPCost = parseInt(val(), "-BP");
        else if (!strcmp(k, "-MGDP")) a.maxBPLog = parseInt(val(), "-MGDP");
        else if (!strcmp(k, "-MNO")) a.OQCMinNonOverlap = parseInt(val(), "-MNO");
        else if (!strcmp(k, "-FBS")) a.FBS = parseBool(val(), "-FBS");
        else if (!strcmp(k, "-PRL")) a.FBS_PSLength = parseFloat(val(), "-PRL");
        else if (!strcmp(k, "-PSS")) a.FBS_PSScore = parseFloat(val(), "-PSS");
        else if (!strcmp(k, "-gpus")) a.gpus = std::max(1, parseInt(val(), "-gpus"));
        else if (!strcmp(k, "-batch")) a.batchReads = std::max(1, parseInt(val(), "-batch"));
        else if (!strcmp(k, "-dev")) a.firstDev = parseInt(val(), "-dev");
        else if (!strcmp(k, "-passes")) a.passes = std::max(1, parseInt(val(), "-passes"));
        else if (!strcmp(k, "-pipes")) a.pipes = std::max(1, parseInt(val(), "-pipes"));
        else if (!strcmp(k, "-replay")) a.replay = true;
        else if (!strcmp(k, "-tpp")) a.threadsPerPipe = parseInt(val(), "-tpp");
        else { fprintf(stderr, "%s is not a valid option.\n\n", k); usage(); exit(1); }
    }
    if (index) {
        if (!a.haveG) { fprintf(stderr, "Genome file specification (-g) is required for index creation.\n\n"); usage(); exit(1); }
        if (a.haveO) { fprintf(stderr, "Output file specification is not allowed during index creation.\n\n"); usage(); exit(1); }
    }
    if (query) {
        if (a.haveG) { fprintf(stderr, "Genome file specification (-g) is not allowed for query alignment.\n"); usage(); exit(1); }
        if (!a.haveX) { fprintf(stderr, "Index file specification (-x) is required for query alignment.\n"); usage(); exit(1); }
        size_t dot = a.xfile.rfind('.');
        if (dot == std::string::npos) { fprintf(stderr, "Specified index filename has improper or missing file extension.  Is it an index file?\n"); exit(1); }
        a.gfile = a.xfile.substr(0, dot) + ".nib2";                              // Main.c:493-501
        if (!a.haveO) { a.outputBlast8 = false; a.outputSAM = true; a.hardClip = true; a.ofile = "stdout"; }
    }
    a.query = query; a.index = index;
    a.postProcess(query);
    return 0;
}

}  // namespace yh
