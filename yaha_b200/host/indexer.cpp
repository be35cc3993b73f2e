// indexer.cpp -- `yaha_b200 -g genome.(fa|nib2) [-L wordLen] [-S skipDist] [-H maxHits]`: the reference's index-creation mode
// (Main.c:567-634) with the index built on the device (ya_open_build).  Writes <stem>.nib2 (from FASTA input; compressFile,
// Compress.c:140-331) and <stem>.X<LL>_<SS>_<HHHHH>S (indexFile, Index.c:49-331; name Main.c:559-563) next to the input, byte
// for byte the files `yaha -g` writes.
#include <errno.h>
#include <string.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "host.hpp"

namespace yh {

static bool writeAll(const std::string &path, const std::vector<std::pair<const void *, size_t>> &parts)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "Failure to open output file: %s.  Error number:%d\n", path.c_str(), errno); return false; }
    for (const auto &p : parts)
        if (p.second && fwrite(p.first, 1, p.second, f) != p.second) { fprintf(stderr, "Failure to write %s\n", path.c_str()); fclose(f); return false; }
    return fclose(f) == 0;
}

// FASTA -> .nib2 v2 image (layout Compress.c:28-63,140-218): header, one 16-byte record per sequence (byte offset of its bases,
// length, name offset, name length), a zero word, the names padded to 4 bytes, then the bases 4 bits each (high nibble first),
// every sequence padded with X to a multiple of 8 bases.  Names are cut at the first blank (Compress.c:277-284).
static bool compressFasta(const std::string &fa, const std::string &nibPath)
{
    FILE *f = fopen(fa.c_str(), "rb");
    if (!f) { fprintf(stderr, "File '%s' does not exist.\n", fa.c_str()); return false; }
    std::string data;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, k);
    fclose(f);
    struct Rec { uint32_t byteOff, length, nameOff, nameLen; };
    std::vector<Rec> recs;
    std::string names;
    std::vector<uint8_t> bases;
    size_t p = data.find('>');
    while (p != std::string::npos && p < data.size()) {
        const size_t nl = data.find('\n', p);
        const size_t next = (nl == std::string::npos) ? std::string::npos : data.find('>', nl);
        std::string name = data.substr(p + 1, (nl == std::string::npos ? data.size() : nl) - (p + 1));
        const size_t sp = name.find(' ');
        if (sp != std::string::npos) name.resize(sp);
        Rec r; r.byteOff = (uint32_t)bases.size(); r.nameOff = (uint32_t)names.size(); r.nameLen = (uint32_t)name.size();
        names += name;
        const size_t b0 = (nl == std::string::npos) ? data.size() : nl + 1, b1 = (next == std::string::npos) ? data.size() : next;
        uint32_t L = 0; int half = -1;
        for (size_t i = b0; i < b1; i++) {
            const char c = data[i];
            if (c == '\n' || c == '\r') continue;
            const int code = codeOfChar((unsigned char)c);
            if (half < 0) half = code; else { bases.push_back((uint8_t)((half << 4) | code)); half = -1; }
            L++;
        }
        uint32_t pad = (8 - (L & 7)) & 7;
        for (uint32_t q = 0; q < pad; q++) { if (half < 0) half = 14; else { bases.push_back((uint8_t)((half << 4) | 14)); half = -1; } }
        r.length = L;
        recs.push_back(r);
        p = next;
    }
    if (recs.empty()) { fprintf(stderr, "No sequences found in %s\n", fa.c_str()); return false; }
    while (names.size() & 3) names.push_back('\0');
    const uint32_t head[4] = {0x01020304u, 2u, (uint32_t)(20 + 16 * recs.size() + names.size()), (uint32_t)recs.size()};
    const uint32_t zero = 0;
    return writeAll(nibPath, {{head, sizeof head}, {recs.data(), recs.size() * sizeof(Rec)}, {&zero, 4}, {names.data(), names.size()}, {bases.data(), bases.size()}});
}

int runIndex(const Args &A)
{
    const size_t dot = A.gfile.rfind('.');
    const std::string stem = (dot == std::string::npos) ? A.gfile : A.gfile.substr(0, dot);
    const std::string ext = (dot == std::string::npos) ? "" : A.gfile.substr(dot);
    std::string nibPath = A.gfile;
    if (ext != ".nib2") {
        nibPath = stem + ".nib2";
        fprintf(stderr, "Compressing %s into %s.\n", A.gfile.c_str(), nibPath.c_str());
        if (!compressFasta(A.gfile, nibPath)) return 1;
    }
    std::string err;
    Genome G;
    if (!G.load(nibPath, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const int K = A.wordLen, S = A.skipDist, H = A.maxHits;
    char name[64];
    snprintf(name, sizeof name, ".X%02d_%02d_%05dS", K, S, H);
    const std::string xPath = stem + name;
    fprintf(stderr, "Creating index file %s.\n", xPath.c_str());
    ya_params P = A.deviceParams();
    P.wordLen = K; P.maxHits = std::min(650, H);
    std::vector<uint32_t> st, ln;
    for (const BaseSeq &b : G.seqs) { st.push_back(b.start); ln.push_back(b.length); }
    ya_ctx *c = ya_open_build(A.firstDev, &P, G.bases, G.nBaseBytes, st.data(), ln.data(), (int)st.size(), (uint32_t)H, (uint32_t)S);
    if (!c) { fprintf(stderr, "yaha_b200: cannot build the index on device %d: %s\n", A.firstDev, ya_last_error(nullptr)); return 1; }
    size_t nSo = 0, nRoa = 0;
    ya_index_sizes(c, &nSo, &nRoa);
    std::vector<uint32_t> so(nSo), roa(nRoa ? nRoa : 1);
    if (ya_index_download(c, so.data(), roa.data()) != YA_OK) { fprintf(stderr, "yaha_b200: %s\n", ya_last_error(c)); ya_close(c); return 1; }
    ya_close(c);
    const uint32_t head[4] = {0xFFFFFFFFu, (uint32_t)K, (uint32_t)H, (uint32_t)nRoa};          // Index.c:171-176
    if (!writeAll(xPath, {{head, sizeof head}, {so.data(), nSo * 4}, {roa.data(), nRoa * 4}})) return 1;
    if (A.verbose) fprintf(stderr, "%zu total %d-mer matches were indexed.\n", nRoa, K);
    return 0;
}

}  // namespace yh
