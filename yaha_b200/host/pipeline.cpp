// pipeline.cpp -- the batched alignment driver: what processQueryFile / processQueries
// (Query.c:255-709) become when the three hot stages run on the device.
//
//   reader thread : parses the query file into batches of reads (input order, bounded queue)
//   pipelines     : P per device, each with its own ya_ctx (sharing the device's resident index),
//                   its own host worker threads and its own fibers.  For one batch:
//                     ya_reads_upload + ya_seed_frags                          (stages 1+2)
//                     one fiber per read runs the reference's per-read control flow and parks in
//                     dpWait(); when all fibers are parked ONE ya_sw_batch executes every posted
//                     job (stage 3); repeat until every fiber has finished
//                   Two pipelines per device overlap one batch's host phase with the other's device
//                   phase.
//   writer        : emits finished batches in input order (== `yaha -t 1` order)
//
// -t N  : host worker threads in total (split over the pipelines)
// -gpus G: G devices, index replicated by peer copy; batches are dealt to whichever pipeline is free,
//          no collective on the data path.
#include <errno.h>
#include <pthread.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/sysinfo.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include "host.hpp"

namespace yh {

static const size_t kStackBytes = 512 * 1024;

// Minimal x86-64 System V context switch: callee-saved registers + stack pointer.  glibc's swapcontext
// makes a signal-mask system call on every switch, which costs more than the rest of a round trip.
extern "C" void yh_switch(void **saveSp, void *loadSp);
__asm__(
    ".text\n.globl yh_switch\n.type yh_switch,@function\n"
    "yh_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size yh_switch,.-yh_switch\n");

struct Slot;
struct Fiber {
    void *sp = nullptr;                   // saved stack pointer while not running
    void *stack = nullptr;
    bool done = false, started = false;
    ReadCtx rc;
    Slot *w = nullptr;
};

// Page-locked result arrays of one ya_sw_batch call.  A block stays alive until every slot served by the
// call has run its next pass (the fibers' DpAnswer views point into it).
struct ResultBlock;

// One worker thread's share of one batch: its fibers, the jobs they posted in the current pass and the
// answers of the previous device call.  A fiber never migrates between OS threads.
struct Slot {
    void *mainSp = nullptr;
    std::vector<Fiber *> fibers;
    int lo = 0, hi = 0;                   // this slot's reads of the batch: [lo, hi)
    std::vector<ya_dp_job> jobs;          // posted in the current pass
    const ya_dp_result *res = nullptr;    // answers (this slot's jobs start at ansBase)
    const ya_op *ops = nullptr;
    size_t ansBase = 0;
    ResultBlock *block = nullptr;         // the block res/ops point into (released after the next pass)
    const Env *E = nullptr;
    bool set = false;                     // fibers created
};

struct Batch;

DpFuture dpSubmit(ReadCtx &rc, int kind, bool rev, uint32_t rOff, int rLen, int qOff, int qLen)
{
    Slot *w = ((Fiber *)rc.owner)->w;
    ya_dp_job j;
    j.rOff = rOff; j.read = (uint32_t)rc.idx; j.rLen = (uint16_t)rLen; j.qOff = (uint16_t)qOff; j.qLen = (uint16_t)qLen;
    j.kind = (uint8_t)kind; j.strand = rev ? 1 : 0;
    w->jobs.push_back(j);
    DpFuture f; f.slot = (int)w->jobs.size() - 1;
    return f;
}

void dpWait(ReadCtx &rc)
{
    if (rc.childSp) { yh_switch(rc.childSp, *rc.parentSp); return; }   // a child fiber parks into its read's fiber
    Fiber *f = (Fiber *)rc.owner;
    yh_switch(&f->sp, f->w->mainSp);
}

DpAnswer dpGet(ReadCtx &rc, DpFuture fu)
{
    Slot *w = ((Fiber *)rc.owner)->w;
    const ya_dp_result &r = w->res[w->ansBase + (size_t)fu.slot];
    DpAnswer a;
    a.score = r.score; a.addedQ = r.addedQLen; a.addedR = r.addedRLen; a.ops = w->ops + r.ops_off; a.n = (int)r.ops_n;
    return a;
}

void dpView(ReadCtx &rc, const ya_dp_result *&res, const ya_op *&ops)
{
    Slot *w = ((Fiber *)rc.owner)->w;
    res = w->res ? w->res + w->ansBase : nullptr;
    ops = w->ops;
}

static std::atomic<uint64_t> gProf[8];
static const bool kProf = getenv("YAHA_B200_PROF") != nullptr;     // phase cycle counters are off unless asked for
static inline uint64_t rdtsc() { unsigned lo, hi; __asm__ volatile("rdtsc" : "=a"(lo), "=d"(hi)); return ((uint64_t)hi << 32) | lo; }
static void readMain(const Env &E, ReadCtx &rc)                       // body of the Query.c:306-497 loop
{
    uint64_t t0 = rdtsc();
    const Args &A = *E.A;
    // (generateRandomSeed, QueryState.c:172-187: the seed is a function of the read alone and its only consumer is the tie
    //  break of the OQC sort, so it is derived there -- seedRandom, oqc.cpp -- for the few reads that get that far)
    for (int rev = 0; rev <= 1; rev++) formClumps(E, rc, rev != 0);
    uint64_t t1 = rdtsc();
    postProcessClumps(E, rc);
    uint64_t t2 = rdtsc();
    if (A.OQC) postFilterBySimilarity(E, rc); else postFilterRemoveDups(E, rc);
    uint64_t t3 = rdtsc();
    formatClumps(E, rc);
    uint64_t t4 = rdtsc();
    for (Clump *c : rc.clumps) delete c;
    rc.clumps.clear();
    if (kProf) { gProf[0] += t1 - t0; gProf[1] += t2 - t1; gProf[2] += t3 - t2; gProf[3] += t4 - t3; gProf[4] += rdtsc() - t4; }
}

static thread_local Fiber *tBoot;
static void fiberEntry()
{
    Fiber *f = tBoot;
    readMain(*f->w->E, f->rc);
    f->done = true;
    yh_switch(&f->sp, f->w->mainSp);
    __builtin_trap();                     // a finished fiber is never resumed
}

static const uint64_t kStackCanary = 0x59414841464942ull;
static inline void checkStack(const void *stack)
{
    if (stack && *(const uint64_t *)stack != kStackCanary) { fprintf(stderr, "yaha_b200: a read's fiber overran its %zu KB stack\n", kStackBytes >> 10); abort(); }
}
struct StackPool {
    std::vector<void *> free_;
    std::mutex mu;
    std::atomic<size_t> made{0};
    void *get()
    {
        { std::lock_guard<std::mutex> g(mu); if (!free_.empty()) { void *p = free_.back(); free_.pop_back(); return p; } }
        // one guard page below the stack: recursion that outgrows it (splitHelper, the OQC quickSort on repeat-rich
        // 32 kb reads) faults instead of silently writing into the neighbouring mapping
        // (a guard page makes every stack two mappings of the process: the first 12 K stacks get one -- far below
        // vm.max_map_count -- the rest, which only a run with tens of thousands of parked reads ever reaches, rely on the canary
        // word at the stack's low end that is checked whenever a fiber leaves the processor)
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const bool guard = made.fetch_add(1) < 12288;
        char *m = (char *)mmap(nullptr, kStackBytes + (guard ? page : 0), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { fprintf(stderr, "cannot allocate fiber stack\n"); exit(1); }
        if (guard) { mprotect(m, page, PROT_NONE); m += page; }
        *(uint64_t *)m = kStackCanary;
        return m;
    }
    void put(void *p) { std::lock_guard<std::mutex> g(mu); free_.push_back(p); }
};
static StackPool gStacks;
// A fiber takes its stack when it first runs and hands it back the moment it finishes, through a per-thread LIFO: most
// reads finish in the pass that starts them (their first DP round was made and answered on the device), so a worker
// keeps running its fibers on the same one or two stacks -- warm in cache and TLB -- instead of touching a fresh
// 512 KB mapping per read.  Only a fiber that parks keeps its stack across passes.
struct StackCache {
    std::vector<void *> v;
    void *get() { if (!v.empty()) { void *p = v.back(); v.pop_back(); return p; } return gStacks.get(); }
    void put(void *p) { if (v.size() < 256) v.push_back(p); else gStacks.put(p); }
    ~StackCache() { for (void *p : v) gStacks.put(p); }
};
static thread_local StackCache tStacks;

// run every unfinished fiber of this slot until it parks or finishes; returns #unfinished
static int slotPass(Slot &w)
{
    int live = 0;
    for (Fiber *f : w.fibers) {
        if (f->done) continue;
        if (!f->started) {
            f->started = true;
            f->stack = tStacks.get();
            // initial frame: six zeroed callee-saved registers, then the entry point as return address;
            // the slot above keeps (%rsp + 8) 16-byte aligned at function entry as the ABI requires
            uintptr_t top = ((uintptr_t)f->stack + kStackBytes) & ~(uintptr_t)15;
            void **sp = (void **)top;
            *--sp = nullptr;
            *--sp = (void *)fiberEntry;
            for (int k = 0; k < 6; k++) *--sp = nullptr;
            f->sp = sp;
            tBoot = f;
        }
        yh_switch(&w.mainSp, f->sp);
        checkStack(f->stack);
        if (!f->done) live++;
        else { tStacks.put(f->stack); f->stack = nullptr; }
    }
    return live;
}


// Growable array in page-locked memory (ya_host_alloc): what the device copies into and out of.
template <class T> struct PinnedVec {
    T *p = nullptr; size_t n = 0, cap = 0;
    PinnedVec() {}
    PinnedVec(const PinnedVec &) = delete;
    PinnedVec &operator=(const PinnedVec &) = delete;
    ~PinnedVec() { ya_host_free(p); }
    T *data() { return p; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
    void resize(size_t want, bool keep = true)   // new elements are not initialised
    {
        if (want > cap) {
            // (never a small block: the batches of handed-back reads are a handful of reads of varying size, and every growth is a
            //  cudaHostAlloc + cudaFreeHost -- milliseconds during which no stream of the process makes progress)
            size_t nc = std::max(std::max(want, cap + cap / 2), ((size_t)256 << 10) / sizeof(T));
            T *q = (T *)ya_host_alloc(nc * sizeof(T));
            if (!q) { fprintf(stderr, "yaha_b200: cannot allocate %zu bytes of page-locked memory\n", nc * sizeof(T)); exit(1); }
            if (n && keep) memcpy(q, p, n * sizeof(T));
            ya_host_free(p);
            p = q; cap = nc;
        }
        n = want;
    }
};

// Page-locked landing place of one batch's device text.  A pipeline owns a small ring of them: pinning memory in the middle
// of a run stalls every stream of the process (YA_ALLOC_LOG), so the ring is filled during the first batches and a
// pipeline whose buffers are all still with the writer waits for one instead of making another.
struct TextBuf {
    PinnedVec<char> text;
    PinnedVec<uint64_t> off;
    PinnedVec<uint8_t> status;
    std::mutex *mu = nullptr; std::condition_variable *cv = nullptr; std::vector<TextBuf *> *home = nullptr;
    void release() { { std::lock_guard<std::mutex> g(*mu); home->push_back(this); } cv->notify_one(); }
};

struct Batch {
    uint64_t seq = 0;
    std::vector<Read> reads;
    std::vector<std::pair<const char *, size_t>> slices;   // FASTA records still to be parsed (by the pipeline that takes the batch)
    std::unique_ptr<Fiber[]> fibers;      // one per read, contiguous (kept when the batch object is recycled)
    int nFibers = 0, fiberCap = 0;
    struct alignas(64) OutBuf { OutText s; };   // (own cache line: every append updates the size)
    std::vector<OutBuf> outBufs;          // formatted records, one buffer per worker thread
    // ya_align_batch: the SAM text of the reads finished on the device (page-locked, lands there straight from the device),
    // where each read's records start, which reads were handed back -- those are run as `residual` through the fibers.
    // The buffers belong to the pipeline that ran the batch (TextBuf ring) and go back to it once the batch is written.
    struct TextBuf *text = nullptr;
    std::unique_ptr<Batch> residual;
    std::vector<int> residualOf;          // residual read k is read residualOf[k] of this batch
    std::vector<std::pair<const char *, size_t>> outRuns;   // what the writer emits for this batch, in input order
    size_t nReads = 0;                    // reads of the batch (they live in `reads`, or only in the pipeline's flat buffers)
};

struct ResultBlock {
    PinnedVec<ya_dp_result> res;
    PinnedVec<ya_op> ops;
    std::atomic<int> users{0};
};

struct Pipe {                              // one batch pipeline: a ya_ctx plus reusable host buffers
    ya_ctx *ctx = nullptr;
    int device = 0;
    PinnedVec<ya_strand_frags> strands;
    PinnedVec<ya_frag> frags;
    PinnedVec<uint32_t> region;
    PinnedVec<uint8_t> codes;
    PinnedVec<uint64_t> offs;
    std::vector<std::unique_ptr<TextBuf>> textBufs;       // ring of device-text landing buffers (at most kTextBufs)
    std::vector<TextBuf *> freeText; std::mutex textMu; std::condition_variable textCv;
    // (the whole ring is pinned when the pipeline sees its first batch, sized for that batch with room to spare: every later
    //  acquisition is a pointer hand-over unless a much larger batch turns up)
    TextBuf *acquireText(size_t nReads, size_t totalBases)
    {
        static const size_t kTextBufs = 3;
        std::unique_lock<std::mutex> lk(textMu);
        if (textBufs.empty()) {
            for (size_t k = 0; k < kTextBufs; k++) {
                textBufs.emplace_back(new TextBuf());
                TextBuf *t = textBufs.back().get();
                t->mu = &textMu; t->cv = &textCv; t->home = &freeText;
                t->off.resize(nReads + 1 + nReads / 2, false); t->status.resize(nReads + nReads / 2 + 1, false);
                t->text.resize(3 * totalBases + 768 * nReads + 4096, false);
                freeText.push_back(t);
            }
        }
        textCv.wait(lk, [&] { return !freeText.empty(); });
        TextBuf *t = freeText.back(); freeText.pop_back();
        return t;
    }
    PinnedVec<char> chars, quals, ids;            // ya_align_batch inputs: the reads as they stand in the file
    PinnedVec<uint32_t> idOffs;
    PinnedVec<uint32_t> clumpFirst, clumpCount;     // ya_form_clumps outputs (clumps of seed fragments made on the device)
    PinnedVec<ya_clump_rec> clumpRecs;
    PinnedVec<ya_frag> clumpPath;
    bool devClumps = false;               // the batch in flight has them
    PinnedVec<ya_prep_rec> prepRecs;      // ya_prepare_clumps outputs (phase 1 of the alignment done on the device)
    PinnedVec<ya_gap_rec> gapRecs;
    PinnedVec<ya_frag> prepPath;
    PinnedVec<ya_dp_job> prepJobs;
    bool devPrep = false;
    ResultBlock *firstBlock = nullptr;    // answers of the device-made jobs: every slot's first pass reads them
    std::vector<ya_dp_job> jobs;
    std::vector<std::unique_ptr<ResultBlock>> blocks;      // every result block this pipeline ever made
    std::vector<ResultBlock *> freeBlocks;
    std::mutex blkMu;
    size_t capJobs = 0, capOps = 0;       // capacity every result block is grown to
    // hand-off between the worker threads and this pipeline's device thread, for the batch in flight
    std::mutex mu;
    std::condition_variable cv;
    std::vector<Slot> slots;              // one per worker thread, reused from batch to batch (vectors keep their capacity)
    std::vector<int> published;           // slots whose pass ended with posted jobs
    int slotsDone = 0, running = 0;       // slots with no unfinished fiber / slots inside a pass
    double tSeed = 0, tDp = 0, tHost = 0, tUpload = 0, tSetup = 0;
    uint64_t nJobs = 0, nRounds = 0;

    ResultBlock *acquireBlock()
    {
        std::lock_guard<std::mutex> g(blkMu);
        if (!freeBlocks.empty()) { ResultBlock *b = freeBlocks.back(); freeBlocks.pop_back(); return b; }
        blocks.emplace_back(new ResultBlock());
        return blocks.back().get();
    }
    void releaseBlock(ResultBlock *b)
    {
        if (b->users.fetch_sub(1) == 1) { std::lock_guard<std::mutex> g(blkMu); freeBlocks.push_back(b); }
    }
};

static void die(ya_ctx *c, const char *what)
{
    fprintf(stderr, "yaha_b200: %s: %s\n", what, ya_last_error(c));
    exit(1);
}

static double nowSec()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// YA_TRACE=<file>: every phase of every thread as one line "kind who batch t0_us t1_us extra" (timeline debugging)
struct TraceEv { char kind; int who; int batch; double t0, t1; int extra; };
static const char *kTracePath = getenv("YA_TRACE");
static std::mutex gTraceMu;
static std::vector<TraceEv> gTrace;
static inline void traceEv(char kind, int who, int batch, double t0, double t1, int extra = 0)
{
    if (!kTracePath) return;
    std::lock_guard<std::mutex> g(gTraceMu);
    gTrace.push_back(TraceEv{kind, who, batch, t0, t1, extra});
}
static void traceDump()
{
    if (!kTracePath || gTrace.empty()) return;
    FILE *f = fopen(kTracePath, "w");
    if (!f) return;
    for (const TraceEv &e : gTrace) fprintf(f, "%c %d %d %.1f %.1f %d\n", e.kind, e.who, e.batch, e.t0 * 1e6, e.t1 * 1e6, e.extra);
    fclose(f);
}

// ----------------------------------------------------------------------------- worker pool
// -t N persistent worker threads shared by every pipeline.  A task is "run one pass over your fibers
// of that batch": worker t always serves slot t, so a fiber stays on one OS thread for its whole life.
struct Task { Pipe *pipe; Batch *batch; Slot *slot; int t; };
struct PoolWorker {
    std::mutex mu; std::condition_variable cv;
    std::deque<Task> q; bool stop = false;
    std::thread th;
};

static void runSlotPass(const Task &t)
{
    Pipe &D = *t.pipe; Slot &s = *t.slot; Batch &B = *t.batch;
    const double h0 = nowSec();
    if (!s.set) {                                                       // first pass: this slot's fibers
        s.set = true;
        s.fibers.reserve((size_t)(s.hi - s.lo));
        size_t bytes = 0;
        for (int i = s.lo; i < s.hi; i++) bytes += B.reads[(size_t)i].fcode.size();
        B.outBufs[(size_t)t.t].s.reserve(2 * bytes + 512 * (size_t)(s.hi - s.lo) + 4096);
        for (int i = s.lo; i < s.hi; i++) {
            Fiber *f = &B.fibers[(size_t)i];
            f->stack = nullptr;                                       // (taken when the fiber first runs, slotPass)
            f->w = &s;
            f->sp = nullptr; f->done = false; f->started = false;     // (the fiber object may be a recycled one)
            f->rc.clumps.clear(); f->rc.primaryCount = 0; f->rc.parked = 0; f->rc.outOff = 0; f->rc.outLen = 0;
            f->rc.owner = f; f->rc.idx = i; f->rc.read = &B.reads[(size_t)i];
            f->rc.out = &B.outBufs[(size_t)t.t].s;
            f->rc.clumps.reserve(8);
            for (int st = 0; st < 2; st++) {
                const ya_strand_frags &sf = D.strands[(size_t)2 * i + st];
                f->rc.frags[st] = D.frags.data() + sf.first;
                f->rc.region[st] = D.region.data() + sf.first;
                f->rc.nFrags[st] = (int)sf.n_frags;
                const size_t seg = (size_t)2 * i + st;
                if (D.devClumps && D.clumpCount[seg] != 0xFFFFFFFFu) {
                    f->rc.devClumps[st] = D.clumpRecs.data() + D.clumpFirst[seg];
                    f->rc.nDevClumps[st] = (int)D.clumpCount[seg];
                    f->rc.devPath[st] = D.devPrep ? D.prepPath.data() : D.clumpPath.data();
                    f->rc.devPrep[st] = D.devPrep ? D.prepRecs.data() + D.clumpFirst[seg] : nullptr;
                    f->rc.devGaps = D.devPrep ? D.gapRecs.data() : nullptr;
                } else { f->rc.devClumps[st] = nullptr; f->rc.nDevClumps[st] = 0; f->rc.devPath[st] = nullptr; f->rc.devPrep[st] = nullptr; }
            }
            s.fibers.push_back(f);
        }
    }
    const int live = slotPass(s);
    if (s.block) { D.releaseBlock(s.block); s.block = nullptr; }      // the answers of the previous call are consumed
    const double dt = nowSec() - h0;
    traceEv('H', t.t, (int)B.seq, h0, h0 + dt, live);
    std::lock_guard<std::mutex> lk(D.mu);
    D.tHost += dt;
    D.running--;
    if (!s.jobs.empty()) D.published.push_back(t.t);
    else {
        if (live != 0) { fprintf(stderr, "yaha_b200: internal error: parked fibers without jobs\n"); exit(1); }
        D.slotsDone++;
    }
    D.cv.notify_one();
}

struct WorkerPool {
    std::vector<std::unique_ptr<PoolWorker>> w;
    void start(int n)
    {
        for (int t = 0; t < n; t++) {
            w.emplace_back(new PoolWorker());
            PoolWorker *me = w.back().get();
            me->th = std::thread([me, t]() {
                if (getenv("YA_PIN")) {                                  // experiment: one worker per core
                    cpu_set_t set; CPU_ZERO(&set); CPU_SET(t % get_nprocs(), &set);
                    pthread_setaffinity_np(pthread_self(), sizeof set, &set);
                }
                for (;;) {
                    Task t;
                    {
                        std::unique_lock<std::mutex> lk(me->mu);
                        me->cv.wait(lk, [&] { return !me->q.empty() || me->stop; });
                        if (me->q.empty()) return;
                        t = me->q.front(); me->q.pop_front();
                    }
                    runSlotPass(t);
                }
            });
        }
    }
    void post(const Task &t)
    {
        PoolWorker &p = *w[(size_t)t.t];
        { std::lock_guard<std::mutex> lk(p.mu); p.q.push_back(t); }
        p.cv.notify_one();
    }
    void stop()
    {
        for (auto &p : w) { { std::lock_guard<std::mutex> lk(p->mu); p->stop = true; } p->cv.notify_one(); }
        for (auto &p : w) p->th.join();
        w.clear();
    }
    int size() const { return (int)w.size(); }
};

// After the first slot of a batch has published its jobs the device thread waits this long (at most)
// for the slots still inside a pass, so that one ya_sw_batch serves many slots.
static int coalesceMicros()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("YA_COALESCE_US"); v = e ? atoi(e) : 200; }
    return v;
}

// Align one batch on one pipeline.  This thread drives the device (stages 1+2, then one ya_sw_batch per
// group of published slots); the shared workers run the fibers.  Results land in each read's rc.out.
// The call-by-call path: stages 1+2 and the DP rounds on the device, everything between them in the reads' fibers.
static void classicPass(const Env &E, Pipe &D, Batch &B, WorkerPool &pool)
{
    double t0 = nowSec();
    const int n = (int)B.reads.size();
    B.nFibers = 0;
    if (n == 0) return;
    D.offs.resize((size_t)n + 1, false);
    size_t total = 0;
    for (int i = 0; i < n; i++) { B.reads[(size_t)i].encode(); D.offs[(size_t)i] = total; total += B.reads[(size_t)i].fcode.size(); }
    D.offs[(size_t)n] = total;
    D.codes.resize(total, false);
    for (int i = 0; i < n; i++) memcpy(D.codes.data() + D.offs[(size_t)i], B.reads[(size_t)i].fcode.data(), B.reads[(size_t)i].fcode.size());
    ya_read_batch rb; rb.n_reads = n; rb.codes = D.codes.data(); rb.offsets = D.offs.data();
    if (ya_reads_upload(D.ctx, &rb) != YA_OK) die(D.ctx, "ya_reads_upload");
    D.tUpload += nowSec() - t0;
    traceEv('U', D.device, (int)B.seq, t0, nowSec());
    t0 = nowSec();
    // stages 1+2
    D.strands.resize((size_t)2 * n, false);
    if (D.frags.size() < (size_t)64 * n) { D.frags.resize((size_t)64 * n, false); D.region.resize((size_t)64 * n, false); }
    ya_frag_batch fb;
    for (;;) {
        fb.frags_cap = D.frags.size(); fb.strands = D.strands.data(); fb.frags = D.frags.data(); fb.region = D.region.data();
        int rcode = ya_seed_frags(D.ctx, &fb);
        if (rcode == YA_E_CAPACITY) { D.frags.resize(fb.frags_needed + 1024, false); D.region.resize(fb.frags_needed + 1024, false); continue; }
        if (rcode != YA_OK) die(D.ctx, "ya_seed_frags");
        break;
    }
    // fragments -> clumps of seed fragments on the device (row N1); strands it leaves out (very many fragments)
    // and everything when YA_HOST_CLUMPS is set are done by the worker that owns the read
    static const bool hostClumps = getenv("YA_HOST_CLUMPS") != nullptr;
    D.devClumps = false;
    if (!hostClumps) {
        const size_t cap = std::max<size_t>(fb.n_frags, 16);
        D.clumpFirst.resize((size_t)2 * n, false); D.clumpCount.resize((size_t)2 * n, false);
        if (D.clumpRecs.size() < cap) { D.clumpRecs.resize(cap + cap / 2, false); D.clumpPath.resize(cap + cap / 2, false); }
        ya_clump_batch cb;
        cb.maxDesert = E.A->maxDesert; cb.minNonOverlap = E.A->minNonOverlap; cb.cap = D.clumpRecs.size();
        cb.clump_first = D.clumpFirst.data(); cb.clump_count = D.clumpCount.data(); cb.clumps = D.clumpRecs.data(); cb.path = D.clumpPath.data();
        const int rcode = ya_form_clumps(D.ctx, &cb);
        if (rcode == YA_OK) D.devClumps = true;
        else if (rcode != YA_E_STATE) die(D.ctx, "ya_form_clumps");            // (YA_E_STATE: survivors not on the device -> host path)
    }
    // ... and the first phase of their alignment (perfect extensions between seed fragments, gap dispatch, extension
    // plan): the jobs come back ready for ya_sw_batch, which this thread runs before any worker has touched the batch
    static const bool hostPrep = getenv("YA_HOST_PREP") != nullptr;
    D.devPrep = false; D.firstBlock = nullptr;
    size_t nPrepJobs = 0;
    if (D.devClumps && !hostPrep) {
        const size_t cap = D.clumpRecs.size();
        if (D.prepRecs.size() < cap) { D.prepRecs.resize(cap, false); D.gapRecs.resize(cap, false); D.prepPath.resize(cap, false); D.prepJobs.resize(3 * cap + 16, false); }
        ya_prep_batch pb;
        pb.cap = cap; pb.jobs_cap = D.prepJobs.size(); pb.prep = D.prepRecs.data(); pb.gaps = D.gapRecs.data(); pb.path = D.prepPath.data();
        pb.jobs = D.prepJobs.data(); pb.n_jobs = 0;
        if (ya_prepare_clumps(D.ctx, &pb) != YA_OK) die(D.ctx, "ya_prepare_clumps");
        D.devPrep = true;
        nPrepJobs = pb.n_jobs;
    }
    D.tSeed += nowSec() - t0;
    traceEv('S', D.device, (int)B.seq, t0, nowSec());

    // one ya_sw_batch for `nj` jobs into a result block with `users` readers
    auto callDp = [&](const ya_dp_job *jobs, int nj, int users) -> ResultBlock * {
        // Every block of a pipeline has the pipeline-wide capacity (monotonic): re-pinning host memory in the
        // middle of a run stalls the whole process (cudaFreeHost / cudaHostAlloc synchronise the device).
        // (~8 jobs per read is typical; doubled when exceeded.  Never below 4096: the batches that come here after
        //  ya_align_batch are the handful of reads it handed back, and a capacity tailored to the first of them would be outgrown --
        //  i.e. re-pinned, a stall of milliseconds for every stream of the process -- by the next slightly larger one)
        if (D.capJobs == 0) D.capJobs = std::max<size_t>((size_t)16 * (size_t)n, 4096);
        if ((size_t)nj > D.capJobs) D.capJobs = 2 * (size_t)nj;
        D.capOps = std::max(D.capOps, 32 * D.capJobs);
        if (D.blocks.empty()) {                                        // first call of this pipeline: pin a few blocks up front
            std::lock_guard<std::mutex> g(D.blkMu);
            for (int k = 0; k < 6; k++) {
                D.blocks.emplace_back(new ResultBlock());
                D.blocks.back()->res.resize(D.capJobs, false);
                D.blocks.back()->ops.resize(D.capOps, false);
                D.freeBlocks.push_back(D.blocks.back().get());
            }
        }
        ResultBlock *blk = D.acquireBlock();
        blk->users.store(users);
        if (blk->res.size() < D.capJobs) blk->res.resize(D.capJobs, false);
        if (blk->ops.size() < D.capOps) blk->ops.resize(D.capOps, false);
        // One device call holds at most 2^31 raw run slots (a job's slots: rows + window + 2, sw.cu): a round of long reads
        // with many clumps each (10 kbp reads over repeats) is served in several calls, answers appended to one block.
        size_t opsUsed = 0;
        for (int lo = 0; lo < nj;) {
            uint64_t slots = 0, rows = 0;                     // (rows bound the back-pointer scratch: <= 768 B per row, 24 GB per call)
            int hi = lo;
            while (hi < nj) {
                const ya_dp_job &j = jobs[hi];
                const uint64_t s = (uint64_t)j.qLen + (j.kind >= YA_DP_EXT_FWD ? (uint64_t)j.qLen + 2u * (uint64_t)E.A->bandWidth : (uint64_t)j.rLen) + 2u;
                if (hi > lo && (slots + s > (1ull << 31) || rows + j.qLen > (1ull << 25))) break;
                slots += s; rows += j.qLen; hi++;
            }
            size_t need = 0;
            int rcode = ya_sw_batch(D.ctx, jobs + lo, hi - lo, blk->res.data() + lo, blk->ops.data() + opsUsed, blk->ops.size() - opsUsed, &need);
            if (rcode == YA_E_CAPACITY) {                   // results are in; only the ops need a bigger buffer
                D.capOps = std::max(D.capOps, 2 * (opsUsed + need) + 1024);
                blk->ops.n = opsUsed;                       // (keep what earlier calls of this round wrote)
                blk->ops.resize(D.capOps, true);
                rcode = ya_sw_fetch_ops(D.ctx, blk->ops.data() + opsUsed, blk->ops.size() - opsUsed);
            }
            if (rcode != YA_OK) die(D.ctx, "ya_sw_batch");
            if (opsUsed) for (int k = lo; k < hi; k++) blk->res[(size_t)k].ops_off += (uint32_t)opsUsed;
            opsUsed += need;
            if (opsUsed >= 0xFFFF0000ull) { fprintf(stderr, "yaha_b200: a DP round returned more than 2^32 edit operations; use a smaller -batch\n"); exit(1); }
            lo = hi;
            D.nRounds++;
        }
        D.nJobs += (uint64_t)nj;
        return blk;
    };

    // fibers: contiguous slices of the batch, one slot per worker thread (set up by the worker itself)
    t0 = nowSec();
    const int nW = pool.size();
    if (B.fiberCap < n) { B.fibers.reset(new Fiber[(size_t)n]); B.fiberCap = n; }
    B.nFibers = n;
    if ((int)B.outBufs.size() != nW) { B.outBufs.clear(); B.outBufs.resize((size_t)nW); }
    for (auto &ob : B.outBufs) ob.s.clear();
    if ((int)D.slots.size() != nW) D.slots = std::vector<Slot>((size_t)nW);
    std::vector<Slot> &slots = D.slots;
    {
        std::lock_guard<std::mutex> lk(D.mu);
        D.published.clear(); D.slotsDone = 0; D.running = 0;
        for (int t = 0; t < nW; t++) {
            Slot &s = slots[(size_t)t];
            s.fibers.clear(); s.jobs.clear(); s.res = nullptr; s.ops = nullptr; s.ansBase = 0; s.block = nullptr; s.set = false;
            s.E = &E;
            s.lo = (int)((int64_t)t * n / nW); s.hi = (int)((int64_t)(t + 1) * n / nW);
            if (s.hi > s.lo) D.running++; else D.slotsDone++;
        }
    }
    D.tSetup += nowSec() - t0;
    if (D.devPrep && nPrepJobs > 0) {                                   // the DP round the device prepared: before any worker runs
        const double h1 = nowSec();
        int users = 0;
        for (int t = 0; t < nW; t++) users += slots[(size_t)t].hi > slots[(size_t)t].lo;
        ResultBlock *blk = callDp(D.prepJobs.data(), (int)nPrepJobs, users);
        for (int t = 0; t < nW; t++) {
            Slot &s = slots[(size_t)t];
            if (s.hi > s.lo) { s.res = blk->res.data(); s.ops = blk->ops.data(); s.ansBase = 0; s.block = blk; }
        }
        D.tDp += nowSec() - h1;
        traceEv('D', users, (int)B.seq, h1, nowSec(), (int)nPrepJobs);
    }
    for (int t = 0; t < nW; t++)
        if (slots[(size_t)t].hi > slots[(size_t)t].lo) pool.post(Task{&D, &B, &slots[(size_t)t], t});

    std::vector<int> take;
    for (;;) {
        take.clear();
        {
            std::unique_lock<std::mutex> lk(D.mu);
            D.cv.wait(lk, [&] { return !D.published.empty() || D.slotsDone == nW; });
            if (D.published.empty()) break;                            // every fiber has finished
            if (D.running > 0 && coalesceMicros() > 0)
                D.cv.wait_for(lk, std::chrono::microseconds(coalesceMicros()), [&] { return D.running == 0; });
            take.swap(D.published);
        }
        const double h1 = nowSec();
        D.jobs.clear();
        for (int t : take) {
            Slot &s = slots[(size_t)t];
            s.ansBase = D.jobs.size();
            D.jobs.insert(D.jobs.end(), s.jobs.begin(), s.jobs.end());
            s.jobs.clear();
        }
        const int nj = (int)D.jobs.size();
        ResultBlock *blk = callDp(D.jobs.data(), nj, (int)take.size());
        D.tDp += nowSec() - h1;
        traceEv('D', (int)take.size(), (int)B.seq, h1, nowSec(), nj);
        {
            std::lock_guard<std::mutex> lk(D.mu);
            D.running += (int)take.size();
        }
        for (int t : take) {
            Slot &s = slots[(size_t)t];
            s.res = blk->res.data(); s.ops = blk->ops.data(); s.block = blk;
            pool.post(Task{&D, &B, &s, t});
        }
    }
    for (int i = 0; i < n; i++) if (B.fibers[(size_t)i].stack) { gStacks.put(B.fibers[(size_t)i].stack); B.fibers[(size_t)i].stack = nullptr; }
}

// The whole per-read path in one device call (ya_align_batch); returns the number of reads handed back.
// FASTA records of a batch parsed straight into the pipeline's page-locked input buffers (no Read objects, no second copy)
static int parseFlat(const Env &E, Pipe &D, Batch &B)
{
    size_t bytes = 0;
    for (const auto &sl : B.slices) bytes += sl.second;
    const size_t cap = B.slices.size();
    D.offs.resize(cap + 1, false); D.idOffs.resize(cap + 1, false);
    D.chars.resize(bytes + 1, false); D.ids.resize(200 * cap + 1, false);
    size_t total = 0, idTotal = 0;
    int n = 0;
    for (const auto &sl : B.slices) {
        D.offs[(size_t)n] = total; D.idOffs[(size_t)n] = (uint32_t)idTotal;
        n += parseFastaRecordInto(sl.first, sl.second, D.chars.data(), &total, D.ids.data(), &idTotal, E.A->maxQueryLength, E.A->wordLen);
    }
    D.offs[(size_t)n] = total; D.idOffs[(size_t)n] = (uint32_t)idTotal;
    B.slices.clear();
    return n;
}

static int fusedPass(const Env &E, Pipe &D, Batch &B, int nFlat)
{
    const double t0 = nowSec();
    const int n = nFlat >= 0 ? nFlat : (int)B.reads.size();
    const bool fastq = E.A->fastq;
    size_t total = 0;
    if (nFlat >= 0) total = (size_t)D.offs[(size_t)n];                 // (parseFlat has filled the buffers)
    else {
        D.offs.resize((size_t)n + 1, false);
        D.idOffs.resize((size_t)n + 1, false);
        size_t idTotal = 0;
        for (int i = 0; i < n; i++) { D.offs[(size_t)i] = total; total += B.reads[(size_t)i].fwd.size(); D.idOffs[(size_t)i] = (uint32_t)idTotal; idTotal += B.reads[(size_t)i].id.size(); }
        D.offs[(size_t)n] = total; D.idOffs[(size_t)n] = (uint32_t)idTotal;
        D.chars.resize(total + 1, false); D.ids.resize(idTotal + 1, false);
        if (fastq) D.quals.resize(total + 1, false);
        for (int i = 0; i < n; i++) {
            const Read &r = B.reads[(size_t)i];
            memcpy(D.chars.data() + D.offs[(size_t)i], r.fwd.data(), r.fwd.size());
            memcpy(D.ids.data() + D.idOffs[(size_t)i], r.id.data(), r.id.size());
            if (fastq) memcpy(D.quals.data() + D.offs[(size_t)i], r.qual.data(), r.qual.size());
        }
    }
    B.text = D.acquireText((size_t)n, total);
    TextBuf &TB = *B.text;
    if (TB.off.size() < (size_t)n + 1) TB.off.resize((size_t)n + 1 + (size_t)n / 4, false);
    if (TB.status.size() < (size_t)n) TB.status.resize((size_t)n + (size_t)n / 4, false);
    if (TB.text.size() < 2 * total + 512 * (size_t)n + 4096) TB.text.resize(3 * total + 768 * (size_t)n + 4096, false);
    ya_text_batch tb;
    memset(&tb, 0, sizeof tb);
    tb.n_reads = n; tb.chars = D.chars.data(); tb.offsets = D.offs.data(); tb.quals = fastq ? D.quals.data() : nullptr;
    tb.ids = D.ids.data(); tb.id_off = D.idOffs.data();
    tb.text = TB.text.data(); tb.text_cap = TB.text.size(); tb.text_off = TB.off.data(); tb.status = TB.status.data();
    D.tUpload += nowSec() - t0;
    const double t1 = nowSec();
    int rcode = ya_align_batch(D.ctx, &tb);
    if (rcode == YA_E_CAPACITY) {
        TB.text.resize(tb.text_needed + tb.text_needed / 4 + 4096, false);
        rcode = ya_align_fetch_text(D.ctx, TB.text.data(), TB.text.size());
    }
    if (rcode == YA_E_STATE) {                        // the batch does not fit one device pass (e.g. > 2^28 seed hits): call-by-call path
        B.text->release(); B.text = nullptr;
        return -1;
    }
    if (rcode != YA_OK) die(D.ctx, "ya_align_batch");
    D.tDp += nowSec() - t1;
    D.nRounds++;
    traceEv('F', D.device, (int)B.seq, t0, nowSec(), tb.n_handed_back);
    return tb.n_handed_back;
}

static bool fusedWanted(const Env &E)
{
    static const bool off = [] { const char *e = getenv("YA_FUSED"); return e && atoi(e) == 0; }();
    static const bool hostSteps = getenv("YA_HOST_CLUMPS") != nullptr || getenv("YA_HOST_PREP") != nullptr;
    return !off && !hostSteps && E.A->outputSAM && !E.A->outputBlast8;
}

// what the writer emits for a batch: every read's records in input order, neighbouring pieces merged into one run
static void addRun(Batch &B, const char *p, size_t len)
{
    if (!len) return;
    if (!B.outRuns.empty() && B.outRuns.back().first + B.outRuns.back().second == p) B.outRuns.back().second += len;
    else B.outRuns.emplace_back(p, len);
}

// Align one batch on one pipeline.  Results land in B.outRuns.
static void processBatch(const Env &E, Pipe &D, Batch &B, WorkerPool &pool)
{
    double t0 = nowSec();
    int nFlat = -1;
    // (-replay keeps the parsed reads of pass 0 as objects; everything else on the ya_align_batch path parses flat)
    if (!B.slices.empty() && fusedWanted(E) && !E.A->replay && B.slices.size() <= 65536) {
        nFlat = parseFlat(E, D, B);
        B.reads.clear();
        traceEv('p', D.device, (int)B.seq, t0, nowSec());
    }
    if (!B.slices.empty()) {                                            // FASTA records cut by the reader, parsed here
        size_t k = 0;
        for (const auto &sl : B.slices) {
            if (k == B.reads.size()) B.reads.emplace_back();            // (Read objects of a recycled batch are refilled in place)
            k += (size_t)parseFastaRecord(sl.first, sl.second, B.reads[k], E.A->maxQueryLength, E.A->wordLen);
        }
        B.reads.resize(k);
        B.slices.clear();
        traceEv('p', D.device, (int)B.seq, t0, nowSec());
    }
    B.outRuns.clear();
    B.nFibers = 0;
    const int n = nFlat >= 0 ? nFlat : (int)B.reads.size();
    B.nReads = (size_t)n;
    if (n == 0) return;
    auto fiberRuns = [](Batch &X, Batch &into) {
        for (int i = 0; i < X.nFibers; i++) { const ReadCtx &rc = X.fibers[(size_t)i].rc; if (rc.outLen) addRun(into, rc.out->data() + rc.outOff, rc.outLen); }
    };
    if (nFlat < 0 && (!fusedWanted(E) || n > 65536)) { classicPass(E, D, B, pool); fiberRuns(B, B); return; }
    const int handed = fusedPass(E, D, B, nFlat);
    if (handed < 0) {                                                   // not as one device pass: every read through the fibers
        if (nFlat >= 0) {
            B.reads.resize((size_t)n);
            for (int i = 0; i < n; i++) {
                Read &r = B.reads[(size_t)i];
                r.id.assign(D.ids.data() + D.idOffs[(size_t)i], D.idOffs[(size_t)i + 1] - D.idOffs[(size_t)i]);
                r.fwd.assign(D.chars.data() + D.offs[(size_t)i], (size_t)(D.offs[(size_t)i + 1] - D.offs[(size_t)i]));
                r.qual.clear(); r.fcode.clear(); r.rcode.clear(); r.rev.clear();
            }
        }
        classicPass(E, D, B, pool);
        fiberRuns(B, B);
        return;
    }
    if (handed == 0) {
        addRun(B, B.text->text.data(), (size_t)B.text->off[(size_t)n]);
        return;
    }
    // the reads the device handed back (a clump to split, a crowded strand, ...) go through the fibers as a batch of their own
    if (!B.residual) B.residual.reset(new Batch());
    Batch &R = *B.residual;
    B.residualOf.clear();
    size_t k = 0;
    for (int i = 0; i < n; i++) {
        if (!B.text->status[(size_t)i]) continue;
        if (k == R.reads.size()) R.reads.emplace_back();
        Read &dst = R.reads[k];
        if (nFlat >= 0) {                                              // (the batch was parsed flat: the read is in the input buffers)
            dst.id.assign(D.ids.data() + D.idOffs[(size_t)i], D.idOffs[(size_t)i + 1] - D.idOffs[(size_t)i]);
            dst.fwd.assign(D.chars.data() + D.offs[(size_t)i], (size_t)(D.offs[(size_t)i + 1] - D.offs[(size_t)i]));
            dst.qual.clear();
        } else {
            const Read &src = B.reads[(size_t)i];
            dst.id = src.id; dst.fwd = src.fwd; dst.qual = src.qual;
        }
        dst.fcode.clear(); dst.rcode.clear(); dst.rev.clear();
        B.residualOf.push_back(i);
        k++;
    }
    R.reads.resize(k);
    R.seq = B.seq;
    classicPass(E, D, R, pool);
    size_t next = 0;                                                      // merge: device text for the finished reads, fiber text for the rest
    for (int i = 0; i < n; i++) {
        if (!B.text->status[(size_t)i]) { addRun(B, B.text->text.data() + B.text->off[(size_t)i], (size_t)(B.text->off[(size_t)i + 1] - B.text->off[(size_t)i])); continue; }
        const ReadCtx &rc = R.fibers[next++].rc;
        if (rc.outLen) addRun(B, rc.out->data() + rc.outOff, rc.outLen);
    }
}

// ----------------------------------------------------------------------------- child fibers of a read
struct ChildFiber { void *sp = nullptr; void *stack = nullptr; bool done = false, started = false; };
struct ChildBoot { ReadCtx *rc; void (*fn)(void *, int); void *arg; int k; ChildFiber *cf; };
static thread_local ChildBoot *tChildBoot;
static void childEntry()
{
    const ChildBoot b = *tChildBoot;                 // (the boot record lives on the parent's stack: copy it first)
    b.fn(b.arg, b.k);
    b.cf->done = true;
    yh_switch(&b.cf->sp, *b.rc->parentSp);
    __builtin_trap();                                // a finished child is never resumed
}

void runAsChildren(ReadCtx &rc, int n, void (*fn)(void *, int), void *arg, std::vector<Clump *> *outs)
{
    PVec<ChildFiber> kids((size_t)n);
    void *parentSp = nullptr;
    rc.parentSp = &parentSp;
    for (;;) {
        int live = 0;
        for (int k = 0; k < n; k++) {
            ChildFiber &c = kids[(size_t)k];
            if (c.done) continue;
            ChildBoot boot{&rc, fn, arg, k, &c};
            if (!c.started) {
                c.started = true;
                c.stack = tStacks.get();
                uintptr_t top = ((uintptr_t)c.stack + kStackBytes) & ~(uintptr_t)15;
                void **sp = (void **)top;
                *--sp = nullptr;
                *--sp = (void *)childEntry;
                for (int q = 0; q < 6; q++) *--sp = nullptr;
                c.sp = sp;
                tChildBoot = &boot;
            }
            rc.clumps.swap(outs[k]);
            rc.childSp = &c.sp;
            yh_switch(&parentSp, c.sp);
            checkStack(c.stack);
            rc.childSp = nullptr;
            rc.clumps.swap(outs[k]);
            if (!c.done) live++;
            else if (c.stack) { tStacks.put(c.stack); c.stack = nullptr; }
        }
        if (live == 0) break;
        dpWait(rc);                                  // one park of the read's fiber serves every parked child
    }
    for (ChildFiber &c : kids) if (c.stack) gStacks.put(c.stack);
    rc.parentSp = nullptr;
}

// bounded, ordered hand-off between reader, pipelines and writer
struct Flow {
    std::mutex mu;
    std::condition_variable cvIn, cvOut, cvSpace;
    std::deque<std::unique_ptr<Batch>> in;
    std::map<uint64_t, std::unique_ptr<Batch>> done;
    bool readerDone = false;
    uint64_t produced = 0;
    size_t maxQueued = 4;
};

int runQueries(const Args &A0)
{
    Args A = A0;
    std::string err;
    Genome G;
    if (!G.load(A.gfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (A.verbose) fprintf(stderr, "Read in %d reference sequences from %s.\n", (int)G.seqs.size(), A.gfile.c_str());
    IndexFile X;
    if (!X.load(A.xfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    A.wordLen = X.wordLen;                                              // Query.c:603-610
    if (X.maxHits < A.maxHits) {
        fprintf(stderr, "WARNING: Index file made with maxHits of %d, while %d specified for this query run.\n"
                "Mimimum of two (%d) will be used.\n", X.maxHits, A.maxHits, X.maxHits);
        A.maxHits = X.maxHits;
    }
    // Queries from standard input (`-q stdin`, the reference's default: Main.c:173-178): the stream cannot be mapped or read
    // twice, so the reader opened here to tell FASTA from FASTQ is the one the reader thread goes on with (sequential reader).
    const bool fromStdin = (A.qfile == "stdin" || A.qfile == "-");
    QueryReader stdinReader;
    if (fromStdin) {
        if (A.passes > 1) { fprintf(stderr, "yaha_b200: -passes needs a query file, not standard input\n"); return 1; }
        if (!stdinReader.open(A.qfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        A.fastq = stdinReader.fastq;
    } else {
        QueryReader probe;                                              // sniff FASTA vs FASTQ before the header is written
        if (!probe.open(A.qfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        A.fastq = probe.fastq;
        probe.close();
    }
    FILE *out = (A.ofile == "stdout") ? stdout : fopen(A.ofile.c_str(), "w");
    if (!out) { fprintf(stderr, "Failure to open output file: %s.  Error number:%d\n", A.ofile.c_str(), errno); return 1; }
    static char obuf[1 << 22];
    setvbuf(out, obuf, _IOFBF, sizeof obuf);
    Env E{&A, &G};

    int nproc = get_nprocs();
    int nThreads = std::max(1, A.numThreads);
    if (nThreads > nproc) {
        fprintf(stderr, "Warning.  Requested number of threads (%d) is greater than number of processors.  %d threads will be used.\n",
                A.numThreads, nproc);
        nThreads = nproc;
    }
    ya_params P = A.deviceParams();
    const int nDev = std::max(1, A.gpus);
    const int pipesPerDev = std::max(1, A.pipes);
    const int nPipes = nDev * pipesPerDev;
    WorkerPool pool;                                                    // -t worker threads, shared by all pipelines
    const int prioLevels = [&] { const char *e = getenv("YA_PRIO"); return std::max(1, std::min(std::min(e ? atoi(e) : 4, 4), pipesPerDev)); }();   // (default: 4 levels; YA_PRIO=0: off)
    pool.start(A.threadsPerPipe > 0 ? std::min(A.threadsPerPipe, 4 * nproc) : nThreads);   // (-tpp N: explicit pool size, may oversubscribe)
    std::vector<Pipe> pipes((size_t)nPipes);
    double tOpen = nowSec();
    double tOpenUpload = 0, tOpenPeer = 0; int peerDirect = 0;
    for (int d = 0; d < nDev; d++) {
        Pipe &first = pipes[(size_t)d * pipesPerDev];
        first.device = A.firstDev + d;
        const double o0 = nowSec();
        first.ctx = (d == 0) ? ya_open(A.firstDev, &P, X.so, X.nSo, X.roa, X.nRoa, G.bases, G.nBaseBytes, G.maxROff)
                             : ya_open_peer(A.firstDev + d, pipes[0].ctx);                 // index replica over NVLink
        if (d == 0) tOpenUpload = nowSec() - o0; else { tOpenPeer += nowSec() - o0; peerDirect += first.ctx ? ya_peer_direct(first.ctx) : 0; }
        if (!first.ctx) { fprintf(stderr, "yaha_b200: cannot open device %d: %s\n", A.firstDev + d, ya_last_error(nullptr)); return 1; }
        for (int k = 1; k < pipesPerDev; k++) {
            Pipe &p = pipes[(size_t)d * pipesPerDev + k];
            p.device = first.device;
            p.ctx = ya_open_shared(first.ctx);
            if (!p.ctx) { fprintf(stderr, "yaha_b200: cannot open shared context: %s\n", ya_last_error(nullptr)); return 1; }
        }
    }
    {
        // what the tail of the per-read path needs on the device (ya_align_batch): output flags and the sequence table
        ya_out_params O;
        O.maxDesert = A.maxDesert; O.minNonOverlap = A.minNonOverlap; O.minRawScore = A.minRawScore; O.minIdentity = A.minIdentity;
        O.OQC = A.OQC; O.FBS = A.FBS; O.OQCMinNonOverlap = A.OQCMinNonOverlap; O.BPCost = A.BPCost; O.maxBPLog = A.maxBPLog;
        O.FBS_PSLength = A.FBS_PSLength; O.FBS_PSScore = A.FBS_PSScore; O.hardClip = A.hardClip; O.fastq = A.fastq;
        std::vector<const char *> names; std::vector<uint32_t> starts, lens;
        for (const BaseSeq &bs : G.seqs) { names.push_back(bs.name.c_str()); starts.push_back(bs.start); lens.push_back(bs.length); }
        for (Pipe &p : pipes)
            if (ya_set_output(p.ctx, &O, (int)names.size(), names.data(), starts.data(), lens.data()) != YA_OK) die(p.ctx, "ya_set_output");
    }
    tOpen = nowSec() - tOpen;
    // YA_START_BARRIER=<dir>:<rank>:<world> (bench.py, one process per GPU): wait here, index resident, until every rank's
    // process is -- otherwise a rank that finishes its 5 GB index upload a moment earlier runs all of its (millisecond) passes
    // while the others still saturate host memory and PCIe with theirs, and the max-over-ranks time measures start-up skew.
    if (const char *sb = getenv("YA_START_BARRIER")) {
        std::string spec(sb);
        const size_t c2 = spec.rfind(':'), c1 = (c2 == std::string::npos) ? c2 : spec.rfind(':', c2 - 1);
        if (c1 != std::string::npos && c2 != std::string::npos) {
            const std::string dir = spec.substr(0, c1);
            const int rank = atoi(spec.substr(c1 + 1, c2 - c1 - 1).c_str()), world = atoi(spec.substr(c2 + 1).c_str());
            FILE *f = fopen((dir + "/ready." + std::to_string(rank)).c_str(), "w");
            if (f) fclose(f);
            for (int waited = 0; waited < 600000; waited++) {            // (ten minutes at most)
                int have = 0;
                for (int r = 0; r < world; r++) have += access((dir + "/ready." + std::to_string(r)).c_str(), F_OK) == 0;
                if (have == world) break;
                usleep(1000);
            }
        }
    }

    std::vector<std::unique_ptr<Batch>> cache;                          // -replay: parsed batches of pass 0
    std::vector<std::unique_ptr<Batch>> spare;                          // written batches, refilled by the reader (strings keep their capacity)
    for (int pass = 0; pass < std::max(1, A.passes); pass++) {
        const bool replaying = A.replay && pass > 0;
        if (pass > 0 && out != stdout) { out = freopen(A.ofile.c_str(), "w", out); setvbuf(out, obuf, _IOFBF, sizeof obuf); }
        writeHeader(E, out);
        for (Pipe &p : pipes) { p.tSeed = p.tDp = p.tHost = p.tUpload = p.tSetup = 0; p.nJobs = p.nRounds = 0; ya_counters c; ya_get_counters(p.ctx, &c); }
        Flow F;
        F.maxQueued = (size_t)nPipes + 2;
        double tRead = 0, tWrite = 0;
        uint64_t nReads = 0;
        std::vector<std::unique_ptr<Batch>> replayIn;
        replayIn.swap(cache);
        const double tStart = nowSec();

        RecordSlicer slicerToClose;
        std::thread reader([&]() {
            uint64_t seq = 0;
            if (replaying) {
                for (auto &b : replayIn) {
                    std::unique_lock<std::mutex> lk(F.mu);
                    F.cvSpace.wait(lk, [&] { return F.in.size() < F.maxQueued; });
                    b->seq = seq++;
                    F.in.push_back(std::move(b));
                    F.cvIn.notify_one();
                }
            } else {
                QueryReader qr;
                RecordSlicer sl;
                std::string e2;
                const bool sliced = !A.fastq && !fromStdin && getenv("YA_SEQ_READER") == nullptr;
                if (fromStdin) { qr = stdinReader; stdinReader.f = nullptr; stdinReader.buf = nullptr; }
                else if (!(sliced ? sl.open(A.qfile, e2) : qr.open(A.qfile, e2))) { fprintf(stderr, "%s\n", e2.c_str()); exit(1); }
                qr.wordLen = A.wordLen; qr.maxLen = A.maxQueryLength;
                sl.wordLen = A.wordLen; sl.maxLen = A.maxQueryLength;
                bool eof = false;
                while (sliced && !eof) {                                  // FASTA: cut records here, parse them in the pipelines
                    double r0 = nowSec();
                    std::unique_ptr<Batch> b;
                    {
                        std::lock_guard<std::mutex> lk(F.mu);
                        if (!spare.empty()) { b = std::move(spare.back()); spare.pop_back(); }
                    }
                    if (!b) { b.reset(new Batch()); b->reads.reserve((size_t)A.batchReads); }
                    b->slices.clear();
                    const char *rs; size_t rn;
                    while ((int)b->slices.size() < A.batchReads) {
                        if (!sl.next(rs, rn)) { eof = true; break; }
                        b->slices.emplace_back(rs, rn);
                    }
                    tRead += nowSec() - r0;
                    traceEv('P', 0, (int)seq, r0, nowSec());
                    if (b->slices.empty()) { std::lock_guard<std::mutex> lk(F.mu); spare.push_back(std::move(b)); break; }
                    b->seq = seq++;
                    std::unique_lock<std::mutex> lk(F.mu);
                    F.cvSpace.wait(lk, [&] { return F.in.size() < F.maxQueued; });
                    F.in.push_back(std::move(b));
                    F.cvIn.notify_one();
                }
                if (sliced) slicerToClose = sl;                             // unmapped after the pipelines have parsed every record
                while (!sliced && !eof) {
                    double r0 = nowSec();
                    std::unique_ptr<Batch> b;
                    {
                        std::lock_guard<std::mutex> lk(F.mu);
                        if (!spare.empty()) { b = std::move(spare.back()); spare.pop_back(); }
                    }
                    if (!b) { b.reset(new Batch()); b->reads.reserve((size_t)A.batchReads); }
                    size_t k = 0;                                       // Read objects of a recycled batch are refilled in place
                    while ((int)k < A.batchReads) {
                        if (k == b->reads.size()) b->reads.emplace_back();
                        if (!qr.next(b->reads[k])) { eof = true; break; }
                        k++;
                    }
                    b->reads.resize(k);
                    tRead += nowSec() - r0;
                    traceEv('P', 0, (int)seq, r0, nowSec());
                    if (b->reads.empty()) break;
                    b->seq = seq++;
                    std::unique_lock<std::mutex> lk(F.mu);
                    F.cvSpace.wait(lk, [&] { return F.in.size() < F.maxQueued; });
                    F.in.push_back(std::move(b));
                    F.cvIn.notify_one();
                }
                if (!sliced) qr.close();
            }
            std::lock_guard<std::mutex> lk(F.mu);
            F.readerDone = true; F.produced = seq;
            F.cvIn.notify_all(); F.cvOut.notify_all();
        });

        std::vector<std::thread> pth;
        for (int p = 0; p < nPipes; p++)
            pth.emplace_back([&, p]() {
                ya_bind_thread(pipes[(size_t)p].ctx);                   // (page-locked buffers of this thread belong to its device)
                for (;;) {
                    std::unique_ptr<Batch> b;
                    {
                        std::unique_lock<std::mutex> lk(F.mu);
                        F.cvIn.wait(lk, [&] { return !F.in.empty() || F.readerDone; });
                        if (F.in.empty()) return;
                        b = std::move(F.in.front());
                        F.in.pop_front();
                        F.cvSpace.notify_one();
                    }
                    // Batches in flight on one device get descending urgency in input order (ya_set_priority): they then finish one
                    // after the other and the ordered writer works on batch k while batch k+1 still computes, instead of four
                    // batches finishing together and their writes queueing behind each other.  YA_PRIO=0: all alike.
                    if (prioLevels > 1) ya_set_priority(pipes[(size_t)p].ctx, (int)((b->seq / (uint64_t)nDev) % (uint64_t)prioLevels));
                    processBatch(E, pipes[(size_t)p], *b, pool);
                    std::lock_guard<std::mutex> lk(F.mu);
                    uint64_t s = b->seq;
                    F.done[s] = std::move(b);
                    F.cvOut.notify_all();
                }
            });

        // writer: input order
        for (uint64_t next = 0;; next++) {
            std::unique_ptr<Batch> b;
            {
                std::unique_lock<std::mutex> lk(F.mu);
                F.cvOut.wait(lk, [&] { return F.done.count(next) || (F.readerDone && next >= F.produced); });
                if (!F.done.count(next)) break;
                b = std::move(F.done[next]);
                F.done.erase(next);
            }
            double w0 = nowSec();
            nReads += b->nReads;
            if (!replaying) {
                // (one ordered stream: writing the batches of a step side by side with pwrite was measured SLOWER -- buffered
                //  writes to one file serialise on its inode lock -- 1.6 ms per 2.3 MB batch with four writers against 0.5 ms)
                for (const auto &run : b->outRuns) fwrite(run.first, 1, run.second, out);
            }
            tWrite += nowSec() - w0;
            traceEv('W', 0, (int)b->seq, w0, nowSec());
            if (b->text) { b->text->release(); b->text = nullptr; }
            if (A.replay) cache.push_back(std::move(b));
            else { std::lock_guard<std::mutex> lk(F.mu); spare.push_back(std::move(b)); }
        }
        reader.join();
        for (auto &t : pth) t.join();
        slicerToClose.close();
        fflush(out);
        const double tAlign = nowSec() - tStart;
        traceEv('A', pass, 0, tStart, tStart + tAlign);
        if (A.verbose || A.passes > 1 || getenv("YAHA_B200_STATS")) {
            ya_counters c{};
            double seed = 0, dp = 0, host = 0, upl = 0, setup = 0; uint64_t jobs = 0, rounds = 0, cells = 0, launches = 0, probes = 0, hits = 0, fragsAll = 0;
            double msdp = 0, msseed = 0, mstb = 0, msext = 0, mslk = 0, msfin = 0; uint64_t extCells = 0, extLaunches = 0, devReads = 0, handed = 0, textBytes = 0;
            for (Pipe &d : pipes) {
                ya_get_counters(d.ctx, &c);
                seed += d.tSeed; dp += d.tDp; host += d.tHost / std::max(1, pool.size()); upl += d.tUpload; setup += d.tSetup; jobs += d.nJobs; rounds += d.nRounds; cells += c.dp_cells;
                msdp += c.ms_dp; msseed += c.ms_seed; mstb += c.ms_traceback; launches += c.launches; probes += c.probes; hits += c.hits;
                fragsAll += c.frags_all; msext += c.ms_ext; mslk += c.ms_lookup; extCells += c.ext_cells; extLaunches += c.ext_launches;
                msfin += c.ms_finish; devReads += c.reads_finished; handed += c.reads_handed_back; textBytes += c.text_bytes;
            }
            // time the bulk extension launches occupied the devices: union of their spans per device (pipelines overlap them)
            double extUnion = 0;
            {
                std::map<int, std::vector<std::pair<float, float>>> byDev;
                std::vector<float> buf;
                for (Pipe &d : pipes) {
                    int np = 0;
                    ya_get_ext_intervals(d.ctx, nullptr, 0, &np);
                    buf.resize((size_t)2 * np + 2);
                    if (np && ya_get_ext_intervals(d.ctx, buf.data(), np, &np) == YA_OK)
                        for (int k = 0; k < np; k++) byDev[d.device].emplace_back(buf[(size_t)2 * k], buf[(size_t)2 * k + 1]);
                }
                for (auto &kv : byDev) {
                    auto &v = kv.second;
                    std::sort(v.begin(), v.end());
                    float hi = -1e30f;
                    for (const auto &iv : v) { if (iv.second <= hi) continue; extUnion += iv.second - std::max(iv.first, hi); hi = iv.second; }
                }
            }
            if (getenv("YA_PASS_CLOCK")) fprintf(stderr, "pass %d started at %.6f (wall clock)\n", pass, std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count() - tAlign);
            fprintf(stderr, "{\"pass\": %d, \"reads\": %llu, \"align_s\": %.5f, \"open_s\": %.3f, \"reads_per_s\": %.1f, \"read_parse_s\": %.5f, "
                    "\"upload_s\": %.5f, \"write_s\": %.5f, \"seed_wall_s\": %.5f, \"dp_wall_s\": %.5f, \"host_wall_s\": %.5f, \"dp_jobs\": %llu, "
                    "\"dp_rounds\": %llu, \"dp_cells\": %llu, \"dev_ms_seed\": %.3f, \"dev_ms_dp\": %.3f, \"dev_ms_traceback\": %.3f, "
                    "\"launches\": %llu, \"probes\": %llu, \"hits\": %llu, \"frags_all\": %llu, \"gpus\": %d, \"threads\": %d, \"pipes\": %d, "
                    "\"replay\": %d, \"dev_ms_ext\": %.4f, \"ext_cells\": %llu, \"ext_launches\": %llu, \"dev_ms_lookup\": %.4f, \"fiber_setup_s\": %.5f, "
                    "\"dev_ms_finish\": %.4f, \"reads_finished_on_device\": %llu, \"reads_handed_back\": %llu, \"device_text_bytes\": %llu, "
                    "\"index_upload_s\": %.3f, \"index_peer_copies_s\": %.3f, \"peer_copies_direct\": %d, \"dev_ms_ext_union\": %.4f}\n",
                    pass, (unsigned long long)nReads, tAlign, tOpen, nReads / std::max(tAlign, 1e-9), tRead, upl, tWrite, seed, dp, host,
                    (unsigned long long)jobs, (unsigned long long)rounds, (unsigned long long)cells, msseed, msdp, mstb,
                    (unsigned long long)launches, (unsigned long long)probes, (unsigned long long)hits, (unsigned long long)fragsAll, nDev, nThreads,
                    nPipes, replaying ? 1 : 0, msext, (unsigned long long)extCells, (unsigned long long)extLaunches, mslk, setup,
                    msfin, (unsigned long long)devReads, (unsigned long long)handed, (unsigned long long)textBytes, tOpenUpload, tOpenPeer, peerDirect, extUnion);
        }
    }
    extern uint64_t gAlignProf[4];
    if (getenv("YAHA_B200_PROF")) fprintf(stderr, "prof align Mcycles: prepare %.1f collapse+ext %.1f apply %.1f score(incl parked) %.1f\n",
                                          gAlignProf[0] / 1e6, gAlignProf[1] / 1e6, gAlignProf[2] / 1e6, gAlignProf[3] / 1e6);
    extern uint64_t gVerdictProf[8];
    if (getenv("YAHA_B200_PROF")) fprintf(stderr, "prof verdicts: reads %llu noclumps %llu nosplit %llu nosplit&<=1scored %llu | clumps scored %llu split %llu drop %llu | reads nosplit&1scored %llu\n",
        (unsigned long long)gVerdictProf[0], (unsigned long long)gVerdictProf[1], (unsigned long long)gVerdictProf[2], (unsigned long long)gVerdictProf[3],
        (unsigned long long)gVerdictProf[4], (unsigned long long)gVerdictProf[5], (unsigned long long)gVerdictProf[6], (unsigned long long)gVerdictProf[7]);
    if (getenv("YAHA_B200_PROF"))
        fprintf(stderr, "prof Mcycles: formClumps %.1f postProcess(incl. parked time) %.1f oqc %.1f format %.1f free %.1f\n", gProf[0] / 1e6,
                gProf[1] / 1e6, gProf[2] / 1e6, gProf[3] / 1e6, gProf[4] / 1e6);
    if (out != stdout) fclose(out); else fflush(out);
    pool.stop();
    traceDump();
    for (int p = nPipes - 1; p >= 0; p--) ya_close(pipes[(size_t)p].ctx);      // shared contexts before their owners
    return 0;
}

}  // namespace yh
