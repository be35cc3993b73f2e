// host.hpp -- host side of yaha_b200: a from-scratch C++ restatement of the parts of yaha 0.1.83
// that north_star keeps on the CPU (fragment graph, clump assembly, score/split, OQC/FBS, SAM
// output, CLI), re-driven in BATCHES so that the three device stages run through the C ABI
// (include/yaha_b200.h) on thousands of reads at a time.
//
// Structure (ours, not the reference's): every read of a batch is a lightweight fiber that runs
// the reference's per-read control flow; whenever it needs DP results it posts jobs and yields;
// when every fiber of the batch is parked the scheduler executes one ya_sw_batch for all posted
// jobs and resumes them.  Semantics follow the reference file:line cited at each function.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <list>
#include "../../include/yaha_b200.h"

namespace yh {

// ----------------------------------------------------------------------------- arguments
struct Args {                       // AlignmentArgs_t, Math.h:257-334; defaults AlignArgs.c:48-87
    std::string gfile, xfile, qfile = "stdin", ofile;
    bool  haveG = false, haveX = false, haveO = false;
    int   numThreads = 1;
    bool  fastq = false;
    int   wordLen = 15, skipDist = 1, maxHits = -1;
    int   bandWidth = 5, maxIntron = -1, minMatch = 25, maxGap = 50, maxDesert = 50, minNonOverlap = -1;
    float minIdentity = 0.9f;
    int   minRawScore = -1;
    bool  affineGapScoring = true;
    int   minExtLength = 0;
    int   GOCost = 5, GECost = 2, RCost = 3, MScore = 1, XCutoff = 25;
    bool  OQC = true;
    int   OQCMinNonOverlap = -1, BPCost = 5, maxBPLog = 5;
    bool  FBS = false;
    float FBS_PSLength = 0.90f, FBS_PSScore = 0.90f;
    int   maxQueryLength = 32000;
    bool  verbose = false, outputBlast8 = false, outputSAM = true, hardClip = true;
    // yaha_b200 extensions (not echoed in @PG)
    int   gpus = 1;                 // -gpus N : shard the read file over N devices
    int   batchReads = 8192;        // -batch N: reads in flight per device
    int   firstDev = 0;             // -dev D  : first CUDA device ordinal
    int   passes = 1;               // -passes N: repeat the whole job N times (bench.py), one stats line per pass
    int   pipes = 2;                // -pipes P : concurrent batch pipelines per device (overlap host and device phases)
    int   threadsPerPipe = 0;       // -tpp N   : size of the shared worker pool when it should differ from -t (may oversubscribe)
    bool  replay = false;           // -replay  : passes after the first reuse the parsed reads and skip the SAM fwrite
    bool  query = false, index = true;
    void  postProcess(bool queryMode);           // AlignArgs.c:108-169
    ya_params deviceParams() const;
};
int parseArgs(int argc, char **argv, Args &a);   // Main.c:187-565 (flags, checks, file names)

// ----------------------------------------------------------------------------- reference genome
struct BaseSeq { std::string name; uint32_t start, length; };       // BaseSequence_t, Math.h:218-225
struct Genome {
    const uint8_t *bases = nullptr;  size_t nBaseBytes = 0;         // packed 4-bit codes
    std::vector<BaseSeq> seqs;
    uint32_t maxROff = 0;                                           // BaseSeq.c:121-125
    const void *map = nullptr; size_t mapLen = 0;
    int code(uint32_t off) const { uint8_t b = bases[off >> 1]; return (off & 1) ? (b & 15) : (b >> 4); }   // Math.c:180-188
    int findSeq(uint32_t off) const;                                // BaseSeq.c:81-90 (-1 if none)
    bool load(const std::string &nib2Path, std::string &err);       // Compress.c:76-134 + BaseSeq.c:115-119
};
struct IndexFile {
    const uint32_t *so = nullptr, *roa = nullptr; size_t nSo = 0, nRoa = 0;
    int wordLen = 0, maxHits = 0;
    const void *map = nullptr; size_t mapLen = 0;
    bool load(const std::string &path, std::string &err);           // Query.c:594-626
};
extern const char kCharOfCode[16];   // Math.c:154
extern const char kCompCharOfCode[16];
extern const uint8_t kCompCode[16];  // Math.c:155
int codeOfChar(int c);               // Math.c:141-152

// ----------------------------------------------------------------------------- reads
struct Read {
    std::string id;                  // <= 200 chars, spaces -> '_' (Query.c:111-135)
    std::string fwd, rev;            // characters, forward and reverse-complement (Query.c:161-168)
    std::string qual;                // FASTQ only
    std::vector<uint8_t> fcode, rcode;
    int len() const { return (int)fwd.size(); }
    void encode();                   // derive fcode from fwd (idempotent)
    void finish();                   // derive rcode / rev from fcode (idempotent)
};
struct QueryReader {                 // readNextQuery, Query.c:102-228
    struct Buf;
    FILE *f = nullptr; Buf *buf = nullptr; bool fastq = false; int maxLen = 32000, wordLen = 15;
    bool returnedOnce = false;                                 // (an empty record is skipped if it is the first one returned, Query.c:304)
    bool open(const std::string &path, std::string &err);     // Query.c:63-74
    bool next(Read &r);                                        // false at EOF
    void close();
};

// FASTA input cut into records without parsing them: the reader thread only finds the '>' markers (a record
// ends at the next marker wherever it stands, Query.c:137-158), the pipeline that takes the batch parses its
// records (parseFastaRecord) -- so parsing runs on as many threads as there are pipelines.
struct RecordSlicer {
    const char *base = nullptr; size_t len = 0, pos = 0; bool done = true;
    int maxLen = 32000, wordLen = 15; bool anyGood = false, skippedEmpty = false;   // for the same rule (set after open)
    bool open(const std::string &path, std::string &err);      // maps the file; the first character is the first marker
    bool next(const char *&s, size_t &n);                      // false at end of input or at an empty record (Query.c:222)
    void close();
};
// 1: r filled; 0: record skipped with the reference's warning (too long / shorter than wordLen / empty)
int parseFastaRecord(const char *s, size_t n, Read &r, int maxLen, int wordLen);
// the same record appended to flat buffers (ids cut to 200 characters): 1 and both lengths advanced, or 0 (skipped, same warnings)
int parseFastaRecordInto(const char *s, size_t n, char *chars, size_t *seqLen, char *ids, size_t *idLen, int maxLen, int wordLen);

// ----------------------------------------------------------------------------- small-block pool
// Per-thread free lists of power-of-two blocks (32 B .. 64 KB), carved from slabs that are never handed back.
// A read makes ~60 short-lived allocations (list nodes of the fragment pieces, clumps, op arrays); a read's fiber
// stays on one worker thread, so get/put never lock and never reach the C library's arena bookkeeping
// (measured: malloc/free/consolidate/trim were ~35 % of the host time before).  -DYH_NO_POOL: plain malloc (ASan).
struct TlsPool {
    static void *get(size_t bytes);
    static void  put(void *p, size_t bytes);
};
template <class T> struct PoolAllocator {
    typedef T value_type;
    PoolAllocator() noexcept {}
    template <class U> PoolAllocator(const PoolAllocator<U> &) noexcept {}
    T *allocate(size_t n) { return (T *)TlsPool::get(n * sizeof(T)); }
    void deallocate(T *p, size_t n) noexcept { TlsPool::put(p, n * sizeof(T)); }
    template <class U> bool operator==(const PoolAllocator<U> &) const noexcept { return true; }
    template <class U> bool operator!=(const PoolAllocator<U> &) const noexcept { return false; }
};

template <class T> using PVec = std::vector<T, PoolAllocator<T>>;      // short-lived per-read arrays

// ----------------------------------------------------------------------------- edit ops
struct Op { uint16_t len; char code; };
// Vector of edit ops with room for a few entries inline: most lists (a seed fragment's single 'M', a
// gap piece) never touch the heap.  Only the std::vector subset the host uses.
class OpVec {
    enum { kInline = 6 };
    Op *p_; uint32_t n_ = 0, cap_ = kInline;
    Op  in_[kInline];
    bool heap() const { return p_ != in_; }
    void grow(size_t want);
public:
    typedef Op *iterator; typedef const Op *const_iterator;
    OpVec() : p_(in_) {}
    OpVec(const OpVec &o) : p_(in_) { assign(o.begin(), o.end()); }
    OpVec(OpVec &&o) noexcept : p_(in_) { swap(o); }
    OpVec &operator=(const OpVec &o) { if (this != &o) assign(o.begin(), o.end()); return *this; }
    OpVec &operator=(OpVec &&o) noexcept { if (this != &o) { clear(); swap(o); } return *this; }
    ~OpVec() { if (heap()) TlsPool::put(p_, (size_t)cap_ * sizeof(Op)); }
    Op *begin() { return p_; }  Op *end() { return p_ + n_; }
    const Op *begin() const { return p_; }  const Op *end() const { return p_ + n_; }
    size_t size() const { return n_; }  bool empty() const { return n_ == 0; }
    Op &operator[](size_t i) { return p_[i]; }  const Op &operator[](size_t i) const { return p_[i]; }
    Op &front() { return p_[0]; }  Op &back() { return p_[n_ - 1]; }
    const Op &front() const { return p_[0]; }  const Op &back() const { return p_[n_ - 1]; }
    void clear() { n_ = 0; }
    void reserve(size_t want) { if (want > cap_) grow(want); }
    void resize(size_t n) { reserve(n); for (size_t i = n_; i < n; i++) p_[i] = Op{0, 0}; n_ = (uint32_t)n; }
    void push_back(Op o) { if (n_ == cap_) grow((size_t)n_ + 1); p_[n_++] = o; }
    void insert(Op *pos, Op o)
    {
        size_t at = (size_t)(pos - p_);
        if (n_ == cap_) grow((size_t)n_ + 1);
        memmove(p_ + at + 1, p_ + at, (n_ - at) * sizeof(Op));
        p_[at] = o; n_++;
    }
    void insert(Op *pos, const Op *first, const Op *last)        // [first,last) must not alias this vector
    {
        size_t at = (size_t)(pos - p_), k = (size_t)(last - first);
        if (k == 0) return;
        reserve(n_ + k);
        memmove(p_ + at + k, p_ + at, (n_ - at) * sizeof(Op));
        memcpy(p_ + at, first, k * sizeof(Op));
        n_ += (uint32_t)k;
    }
    void erase(Op *pos) { erase(pos, pos + 1); }
    void erase(Op *first, Op *last)
    {
        memmove(first, last, (size_t)(end() - last) * sizeof(Op));
        n_ -= (uint32_t)(last - first);
    }
    void assign(const Op *first, const Op *last)
    {
        size_t k = (size_t)(last - first);
        n_ = 0; reserve(k);
        memcpy(p_, first, k * sizeof(Op)); n_ = (uint32_t)k;
    }
    void swap(OpVec &o) noexcept;
    // filled from outside: room for `cap` entries, then the number actually written
    Op *fill(size_t cap) { n_ = 0; reserve(cap); return p_; }
    void setSize(size_t n) { n_ = (uint32_t)n; }
};
struct OpList {                      // EditOpList_t semantics, SW.cpp:114-283, SW.inl:66-78
    OpVec v;
    bool empty() const { return v.empty(); }
    void clear() { v.clear(); }
    void pushFront(char c, int len) { v.insert(v.begin(), Op{(uint16_t)len, c}); }    // no coalescing
    void pushBack(char c, int len) { v.push_back(Op{(uint16_t)len, c}); }
    void mergeToFront(OpList &src);  // this = src + this, coalescing the junction (SW.cpp:151-205)
    void mergeToBack(OpList &src);   // this = this + src, coalescing the junction (SW.cpp:207-261)
    void mergeToFront(const ya_op *o, int n);   // same, from a device result range
    void mergeToBack(const ya_op *o, int n);
};

// ----------------------------------------------------------------------------- clumps
typedef ya_frag Frag;                // Fragment_t, Math.h:448-455
inline int      fragQLen(const Frag &f) { return 1 + (int)f.endQueryOff - (int)f.startQueryOff; }
inline uint32_t fragERO(const Frag &f) { return f.startRefOff + f.refLen - 1; }
inline void     fragSetERO(Frag &f, uint32_t ro) { f.refLen = (uint16_t)(1 + ro - f.startRefOff); }
inline uint32_t fragDiag(const Frag &f) { return f.startRefOff - f.startQueryOff; }
inline uint32_t absDiff(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }     // FragsClumps.inl:133-137
inline unsigned calcGap(int lo, int hi) { return hi > lo ? (unsigned)(hi - lo) - 1 : 0; }           // :157
inline unsigned calcOverlap(int lo, int hi) { return lo >= hi ? (unsigned)(lo - hi) + 1 : 0; }      // :158
inline unsigned calcGapU(uint32_t lo, uint32_t hi) { return hi > lo ? (hi - lo) - 1 : 0; }
inline unsigned calcOverlapU(uint32_t lo, uint32_t hi) { return lo >= hi ? (lo - hi) + 1 : 0; }

struct SFrag { Frag frag; int score = 0; OpList ops; };              // SFragment_t, Math.h:469-477
typedef std::list<SFrag, PoolAllocator<SFrag>> SFragList;
enum { kReversed = 1, kFormed = 2, kAligned = 4, kScored = 8, kSplit = 16, kPrimary = 32 };   // FragsClumps.inl:221-226
struct Clump {                       // Clump_t, Math.h:511-527
    OpList ops;
    PVec<Frag> path;                 // before alignment: the seed fragments of the clump in query order (formClumps)
    const ya_prep_rec *prep = nullptr;   // phase 1 of the alignment done on the device (ya_prepare_clumps): plan + gaps
    const ya_gap_rec *gapBase = nullptr;
    SFragList sf;                    // after alignment: the collapsed piece (and what splitting makes of it)
    static void *operator new(size_t n) { return TlsPool::get(n); }
    static void operator delete(void *p, size_t n) { TlsPool::put(p, n); }
    uint16_t totScore = 0, totLength = 0, matchedBases = 0, mismatchedBases = 0, gapBases = 0;
    uint16_t numSecondaries = 0, matchedPrimary = 0;
    uint8_t  status = 0, mapQuality = 255;
    bool is(int flag) const { return (status & flag) != 0; }
    void set(int flag, bool on) { if (on) status |= flag; else status &= ~flag; }
    bool reversed() const { return is(kReversed); }
    const Frag &firstFrag() const { return sf.empty() ? path.front() : sf.front().frag; }
    const Frag &lastFrag() const { return sf.empty() ? path.back() : sf.back().frag; }
    uint16_t SQO() const { return firstFrag().startQueryOff; }
    uint16_t EQO() const { return lastFrag().endQueryOff; }
    uint32_t SRO() const { return firstFrag().startRefOff; }
    uint32_t ERO() const { return fragERO(lastFrag()); }
};

struct RandState { uint32_t s[5]; uint32_t bits(); };               // Math.c:274-284

// ----------------------------------------------------------------------------- per-read state
struct DpFuture { int slot = -1; };  // index into the round's job list
struct DpAnswer { int score = 0; int addedQ = 0, addedR = 0; const ya_op *ops = nullptr; int n = 0; };   // view into the round's result arrays

// Growable text buffer of one worker's formatted records.  Unlike std::string it hands out uninitialised room: a SAM record
// is written through a raw cursor into space reserved for its largest possible size (a few KB), and value-initialising that
// space for every record cost about a quarter of the formatting time.
struct OutText {
    char *p = nullptr; size_t n = 0, cap = 0;
    OutText() {}
    OutText(OutText &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    OutText &operator=(OutText &&o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
    OutText(const OutText &) = delete;
    OutText &operator=(const OutText &) = delete;
    ~OutText() { free(p); }
    size_t size() const { return n; }
    const char *data() const { return p; }
    void clear() { n = 0; }
    void reserve(size_t want)
    {
        if (want <= cap) return;
        size_t c = cap ? cap : 4096;
        while (c < want) c *= 2;
        char *q = (char *)realloc(p, c);
        if (!q) { fprintf(stderr, "yaha_b200: out of memory\n"); abort(); }
        p = q; cap = c;
    }
    char *room(size_t k) { reserve(n + k); return p + n; }          // k writable bytes behind the text; commit() what was used
    void commit(size_t k) { n += k; }
    void append(const char *s, size_t k) { memcpy(room(k), s, k); n += k; }
};

struct ReadCtx {                     // the per-read half of QueryState_t (Math.h:587-666)
    void  *owner = nullptr;         // the fiber running this read (scheduler private)
    int    idx = 0;                  // index in the batch == read id on the device
    Read  *read = nullptr;
    std::vector<Clump *> clumps;     // LIFO list: back() is the reference's list head
    std::vector<Clump *> scratch;    // second list of the same kind (filters build their result here and swap); keeps its capacity
    int    primaryCount = 0;
    RandState rng;
    uint64_t parked = 0;             // cycles spent parked in the scoring phase (YAHA_B200_PROF only)
    Frag  *frags[2] = {nullptr, nullptr};        // the device's surviving fragments of each strand, in the
    const uint32_t *region[2] = {nullptr, nullptr};  // pipeline's result buffers (edited in place)
    int    nFrags[2] = {0, 0};
    const ya_clump_rec *devClumps[2] = {nullptr, nullptr};   // clumps formed on the device (ya_form_clumps), or null
    const Frag *devPath[2] = {nullptr, nullptr};
    int    nDevClumps[2] = {0, 0};
    void **childSp = nullptr;        // set while a child fiber of this read runs (runAsChildren): where dpWait saves it ...
    void **parentSp = nullptr;       // ... and the read's own context it returns to
    const ya_prep_rec *devPrep[2] = {nullptr, nullptr};             // ... and prepared there (ya_prepare_clumps), or null
    const ya_gap_rec *devGaps = nullptr;
    OutText *out = nullptr;          // the worker's output buffer; this read's records are [outOff, outOff+outLen)
    size_t outOff = 0, outLen = 0;
    // the reverse-complement strand is derived the first time something asks for it (about half of the reads never do)
    const uint8_t *codes(bool rev) const { if (rev) read->finish(); return rev ? read->rcode.data() : read->fcode.data(); }
    const std::string &chars(bool rev) const { if (rev) read->finish(); return rev ? read->rev : read->fwd; }
};

// Implemented by the scheduler (pipeline.cpp); callable from inside a read fiber.
DpFuture dpSubmit(ReadCtx &rc, int kind, bool rev, uint32_t rOff, int rLen, int qOff, int qLen);
void     dpWait(ReadCtx &rc);                      // park until the round's jobs are done
DpAnswer dpGet(ReadCtx &rc, DpFuture f);        // valid until this fiber parks again
void     dpView(ReadCtx &rc, const ya_dp_result *&res, const ya_op *&ops);   // the round's arrays: res[slot], ops + res[slot].ops_off
// Runs fn(arg, k), k = 0..n-1, as child fibers of the read's fiber: each runs until it parks in dpWait or returns;
// when all live children are parked the read's fiber parks ONCE for all of them, so their DP requests share a round.
// While child k runs, rc.clumps is outs[k] (what it appends to the read's clump list lands there).
void runAsChildren(ReadCtx &rc, int n, void (*fn)(void *arg, int k), void *arg, std::vector<Clump *> *outs);

// ----------------------------------------------------------------------------- algorithms
struct Env { const Args *A; const Genome *G; };
// graph.cpp: fragments -> clumps (QueryMatch.c:224-303, GraphPath.cpp:57-292, AlignHelpers.c:48-193)
void formClumps(const Env &E, ReadCtx &rc, bool rev);
// align.cpp: AlignHelpers.c:205-579, AlignExtFrag.cpp:30-234, careful trimming SW.cpp:553-788
void postProcessClumps(const Env &E, ReadCtx &rc);
// oqc.cpp: GraphPath.cpp:294-1174
void postFilterBySimilarity(const Env &E, ReadCtx &rc);
void postFilterRemoveDups(const Env &E, ReadCtx &rc);
// sam.cpp: AlignOutput.c:30-321
void writeHeader(const Env &E, FILE *out);
void formatClumps(const Env &E, ReadCtx &rc);

// pipeline.cpp
int runQueries(const Args &A);      // processQueryFile, Query.c:551-709, batched
int runIndex(const Args &A);        // indexFile / compressFile front end, Main.c:567-634
void disposeClump(Clump *c);

}  // namespace yh
