// pipeline.cpp -- the batched alignment driver: what processQueryFile / processQueries
// (Query.c:255-709) become when the three hot stages run on the device.
//
//   reader  : fills a batch of reads (sequential, input order)
//   device  : ya_reads_upload + ya_seed_frags for the whole batch          (stages 1+2)
//   fibers  : one per read, running the reference's per-read control flow; each parks in dpWait()
//   rounds  : when all fibers of the batch are parked, ONE ya_sw_batch executes every posted job
//             (stage 3), answers are distributed and the fibers resume
//   writer  : emits the per-read record buffers in input order (== `yaha -t 1` order)
//
// -t N gives N host worker threads per device, each owning a slice of the batch's fibers.
// -gpus G shards the read stream over G devices in contiguous blocks (replicated index, no
// collective on the data path); blocks are written in input order.
#include <pthread.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/sysinfo.h>
#include <ucontext.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include "host.hpp"

namespace yh {

static const size_t kStackBytes = 512 * 1024;

struct Worker;
struct Fiber {
    ucontext_t ctx;
    void *stack = nullptr;
    bool done = false, started = false;
    ReadCtx rc;
    Worker *w = nullptr;
};

struct Worker {
    ucontext_t mainCtx;
    std::vector<Fiber *> fibers;
    std::vector<ya_dp_job> jobs;          // posted in the current round
    std::vector<DpAnswer> answers;        // results of the previous round's jobs
    Fiber *cur = nullptr;
    const Env *E = nullptr;
    size_t jobBase = 0;                   // offset of this worker's jobs in the merged round list
};

struct Batch {
    std::vector<Read> reads;
    std::vector<std::unique_ptr<Fiber>> fibers;
};

DpFuture dpSubmit(ReadCtx &rc, int kind, bool rev, uint32_t rOff, int rLen, int qOff, int qLen)
{
    Worker *w = ((Fiber *)rc.owner)->w;
    ya_dp_job j;
    j.rOff = rOff; j.read = (uint32_t)rc.idx; j.rLen = (uint16_t)rLen; j.qOff = (uint16_t)qOff; j.qLen = (uint16_t)qLen;
    j.kind = (uint8_t)kind; j.strand = rev ? 1 : 0;
    w->jobs.push_back(j);
    DpFuture f; f.slot = (int)w->jobs.size() - 1;
    return f;
}

void dpWait(ReadCtx &rc)
{
    Fiber *f = (Fiber *)rc.owner;
    swapcontext(&f->ctx, &f->w->mainCtx);
}

DpAnswer &dpGet(ReadCtx &rc, DpFuture fu)
{
    Worker *w = ((Fiber *)rc.owner)->w;
    return w->answers[(size_t)fu.slot];
}

static void readMain(const Env &E, ReadCtx &rc)                       // body of the Query.c:306-497 loop
{
    const Args &A = *E.A;
    // generateRandomSeed, QueryState.c:172-187
    {
        const std::vector<uint8_t> &c = rc.read->fcode;
        size_t q = 0;
        for (int i = 0; i < 5; i++) {
            uint32_t word = 0;
            for (int j = 0; j < 16; j++) { word = (word << 2) | (c[q] & 3u); if (++q >= c.size()) q = 0; }
            rc.rng.s[i] = word;
        }
    }
    for (int rev = 0; rev <= 1; rev++) formClumps(E, rc, rev != 0);
    postProcessClumps(E, rc);
    if (A.OQC) postFilterBySimilarity(E, rc); else postFilterRemoveDups(E, rc);
    formatClumps(E, rc);
    for (Clump *c : rc.clumps) delete c;
    rc.clumps.clear();
}

static void fiberEntry(unsigned lo, unsigned hi)
{
    Fiber *f = (Fiber *)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    readMain(*f->w->E, f->rc);
    f->done = true;
    swapcontext(&f->ctx, &f->w->mainCtx);
}

// run every unfinished fiber of this worker until it parks or finishes; returns #unfinished
static int workerRound(Worker &w)
{
    int live = 0;
    for (Fiber *f : w.fibers) {
        if (f->done) continue;
        w.cur = f;
        if (!f->started) {
            f->started = true;
            getcontext(&f->ctx);
            f->ctx.uc_stack.ss_sp = f->stack;
            f->ctx.uc_stack.ss_size = kStackBytes;
            f->ctx.uc_link = &w.mainCtx;
            uintptr_t p = (uintptr_t)f;
            makecontext(&f->ctx, (void (*)())fiberEntry, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
        }
        swapcontext(&w.mainCtx, &f->ctx);
        if (!f->done) live++;
    }
    return live;
}

struct StackPool {
    std::vector<void *> free_;
    std::mutex mu;
    void *get()
    {
        { std::lock_guard<std::mutex> g(mu); if (!free_.empty()) { void *p = free_.back(); free_.pop_back(); return p; } }
        void *p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { fprintf(stderr, "cannot allocate fiber stack\n"); exit(1); }
        return p;
    }
    void put(void *p) { std::lock_guard<std::mutex> g(mu); free_.push_back(p); }
};
static StackPool gStacks;

struct Device {
    ya_ctx *ctx = nullptr;
    int ordinal = 0;
    // stage outputs
    std::vector<ya_strand_frags> strands;
    std::vector<ya_frag> frags;
    std::vector<uint32_t> region;
    std::vector<uint8_t> codes;
    std::vector<uint64_t> offs;
    std::vector<ya_dp_job> jobs;
    std::vector<ya_dp_result> res;
    std::vector<ya_op> ops;
    double tSeed = 0, tDp = 0, tHost = 0, tUpload = 0;
    uint64_t nJobs = 0, nRounds = 0;
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

// Align one batch on one device with `nThreads` host workers.  Results land in each read's rc.out.
static void processBatch(const Env &E, Device &D, Batch &B, int nThreads)
{
    const int n = (int)B.reads.size();
    if (n == 0) return;
    double t0 = nowSec();
    // upload forward codes
    D.offs.resize((size_t)n + 1);
    size_t total = 0;
    for (int i = 0; i < n; i++) { D.offs[(size_t)i] = total; total += B.reads[(size_t)i].fcode.size(); }
    D.offs[(size_t)n] = total;
    D.codes.resize(total);
    for (int i = 0; i < n; i++) memcpy(D.codes.data() + D.offs[(size_t)i], B.reads[(size_t)i].fcode.data(), B.reads[(size_t)i].fcode.size());
    ya_read_batch rb; rb.n_reads = n; rb.codes = D.codes.data(); rb.offsets = D.offs.data();
    if (ya_reads_upload(D.ctx, &rb) != YA_OK) die(D.ctx, "ya_reads_upload");
    D.tUpload += nowSec() - t0;
    t0 = nowSec();
    // stages 1+2
    D.strands.resize((size_t)2 * n);
    if (D.frags.size() < (size_t)64 * n) { D.frags.resize((size_t)64 * n); D.region.resize((size_t)64 * n); }
    ya_frag_batch fb;
    for (;;) {
        fb.frags_cap = D.frags.size(); fb.strands = D.strands.data(); fb.frags = D.frags.data(); fb.region = D.region.data();
        int rcode = ya_seed_frags(D.ctx, &fb);
        if (rcode == YA_E_CAPACITY) { D.frags.resize(fb.frags_needed + 1024); D.region.resize(fb.frags_needed + 1024); continue; }
        if (rcode != YA_OK) die(D.ctx, "ya_seed_frags");
        break;
    }
    double t1 = nowSec();
    D.tSeed += t1 - t0;

    // fibers
    B.fibers.clear();
    B.fibers.reserve((size_t)n);
    std::vector<Worker> workers((size_t)nThreads);
    for (int t = 0; t < nThreads; t++) workers[(size_t)t].E = &E;
    for (int i = 0; i < n; i++) {
        std::unique_ptr<Fiber> f(new Fiber());
        f->stack = gStacks.get();
        Worker &w = workers[(size_t)((int64_t)i * nThreads / n)];
        f->w = &w;
        f->rc.owner = f.get(); f->rc.idx = i; f->rc.read = &B.reads[(size_t)i];
        for (int st = 0; st < 2; st++) {
            const ya_strand_frags &s = D.strands[(size_t)2 * i + st];
            f->rc.frags[st].assign(D.frags.begin() + s.first, D.frags.begin() + s.first + s.n_frags);
            f->rc.region[st].assign(D.region.begin() + s.first, D.region.begin() + s.first + s.n_frags);
        }
        w.fibers.push_back(f.get());
        B.fibers.push_back(std::move(f));
    }

    // rounds: every worker thread keeps its own fibers for the whole batch (a fiber never migrates
    // between OS threads); thread 0 runs the device call between two barriers
    std::atomic<int> live(0);
    bool finished = false;
    auto deviceRound = [&]() {
        double h1 = nowSec();
        D.jobs.clear();
        for (Worker &w : workers) { w.jobBase = D.jobs.size(); D.jobs.insert(D.jobs.end(), w.jobs.begin(), w.jobs.end()); }
        if (D.jobs.empty()) {
            if (live.load() != 0) { fprintf(stderr, "yaha_b200: internal error: parked fibers without jobs\n"); exit(1); }
            finished = true;
            return;
        }
        const int nj = (int)D.jobs.size();
        D.res.resize((size_t)nj);
        if (D.ops.size() < (size_t)nj * 8) D.ops.resize((size_t)nj * 8);
        for (;;) {
            size_t need = 0;
            int rcode = ya_sw_batch(D.ctx, D.jobs.data(), nj, D.res.data(), D.ops.data(), D.ops.size(), &need);
            if (rcode == YA_E_CAPACITY) { D.ops.resize(need + 1024); continue; }
            if (rcode != YA_OK) die(D.ctx, "ya_sw_batch");
            break;
        }
        D.nJobs += (uint64_t)nj; D.nRounds++;
        for (Worker &w : workers) {
            const size_t m = w.jobs.size();
            w.answers.clear();
            w.answers.resize(m);
            for (size_t k = 0; k < m; k++) {
                const ya_dp_result &r = D.res[w.jobBase + k];
                DpAnswer &a = w.answers[k];
                a.score = r.score; a.addedQ = r.addedQLen; a.addedR = r.addedRLen;
                a.ops.v.resize(r.ops_n);
                for (uint32_t q = 0; q < r.ops_n; q++) { const ya_op &o = D.ops[r.ops_off + q]; a.ops.v[q] = Op{o.length, (char)o.opcode}; }
            }
            w.jobs.clear();
        }
        live.store(0);
        D.tDp += nowSec() - h1;
    };
    if (nThreads == 1) {
        while (!finished) {
            double h0 = nowSec();
            live += workerRound(workers[0]);
            D.tHost += nowSec() - h0;
            deviceRound();
        }
    } else {
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, (unsigned)nThreads);
        std::vector<std::thread> th;
        for (int t = 0; t < nThreads; t++)
            th.emplace_back([&, t]() {
                for (;;) {
                    double h0 = nowSec();
                    live += workerRound(workers[(size_t)t]);
                    pthread_barrier_wait(&bar);
                    if (t == 0) { D.tHost += nowSec() - h0; deviceRound(); }
                    pthread_barrier_wait(&bar);
                    if (finished) break;
                }
            });
        for (auto &x : th) x.join();
        pthread_barrier_destroy(&bar);
    }
    for (auto &f : B.fibers) { gStacks.put(f->stack); f->stack = nullptr; }
}

int runQueries(const Args &A0)
{
    Args A = A0;
    std::string err;
    QueryReader qr;
    if (!qr.open(A.qfile == "stdout" ? "stdin" : A.qfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    A.fastq = qr.fastq;
    Genome G;
    if (!G.load(A.gfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (A.verbose) fprintf(stderr, "Read in %d reference sequences from %s.\n", (int)G.seqs.size(), A.gfile.c_str());
    FILE *out = (A.ofile == "stdout") ? stdout : fopen(A.ofile.c_str(), "w");
    if (!out) { fprintf(stderr, "Failure to open output file: %s.  Error number:%d\n", A.ofile.c_str(), errno); return 1; }
    static char obuf[1 << 22];
    setvbuf(out, obuf, _IOFBF, sizeof obuf);
    IndexFile X;
    if (!X.load(A.xfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    A.wordLen = X.wordLen;                                              // Query.c:603-610
    if (X.maxHits < A.maxHits) {
        fprintf(stderr, "WARNING: Index file made with maxHits of %d, while %d specified for this query run.\n"
                "Mimimum of two (%d) will be used.\n", X.maxHits, A.maxHits, X.maxHits);
        A.maxHits = X.maxHits;
    }
    Env E{&A, &G};
    writeHeader(E, out);
    qr.wordLen = A.wordLen; qr.maxLen = A.maxQueryLength;

    int nproc = get_nprocs();
    int nThreads = std::max(1, A.numThreads);
    if (nThreads > nproc) {
        fprintf(stderr, "Warning.  Requested number of threads (%d) is greater than number of processors.  %d threads will be used.\n",
                A.numThreads, nproc);
        nThreads = nproc;
    }
    ya_params P = A.deviceParams();
    const int nDev = std::max(1, A.gpus);
    std::vector<Device> devs((size_t)nDev);
    double tOpen = nowSec();
    for (int d = 0; d < nDev; d++) {
        devs[(size_t)d].ordinal = d;
        devs[(size_t)d].ctx = (d == 0) ? ya_open(A.firstDev, &P, X.so, X.nSo, X.roa, X.nRoa, G.bases, G.nBaseBytes, G.maxROff)
                                       : ya_open_peer(A.firstDev + d, devs[0].ctx);   // index replica over NVLink
        if (!devs[(size_t)d].ctx) { fprintf(stderr, "yaha_b200: cannot open device %d: %s\n", d, ya_last_error(nullptr)); return 1; }
    }
    tOpen = nowSec() - tOpen;

    // Reads are dealt to devices in contiguous blocks of batchReads; blocks are written in order.
    const int threadsPerDev = std::max(1, nThreads / nDev);
    for (int pass = 0; pass < std::max(1, A.passes); pass++) {
        if (pass > 0) {                                                 // -passes N (bench): redo the whole job
            qr.close();
            if (!qr.open(A.qfile == "stdout" ? "stdin" : A.qfile, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
            qr.wordLen = A.wordLen; qr.maxLen = A.maxQueryLength;
            if (out != stdout) { out = freopen(A.ofile.c_str(), "w", out); setvbuf(out, obuf, _IOFBF, sizeof obuf); }
            writeHeader(E, out);
        }
        for (Device &d : devs) { d.tSeed = d.tDp = d.tHost = d.tUpload = 0; d.nJobs = d.nRounds = 0; ya_counters c; ya_get_counters(d.ctx, &c); }
        uint64_t nReads = 0;
        double tRead = 0, tWrite = 0;
        double tAlign = nowSec();
        bool eof = false;
        while (!eof) {
            std::vector<Batch> batches((size_t)nDev);
            int used = 0;
            double r0 = nowSec();
            for (int d = 0; d < nDev && !eof; d++) {
                Batch &B = batches[(size_t)d];
                B.reads.reserve((size_t)A.batchReads);
                while ((int)B.reads.size() < A.batchReads) {
                    B.reads.emplace_back();
                    if (!qr.next(B.reads.back())) { B.reads.pop_back(); eof = true; break; }
                }
                if (!B.reads.empty()) used = d + 1;
            }
            tRead += nowSec() - r0;
            if (used == 0) break;
            if (used == 1) processBatch(E, devs[0], batches[0], nThreads);
            else {
                std::vector<std::thread> th;
                for (int d = 0; d < used; d++) th.emplace_back([&, d]() { processBatch(E, devs[(size_t)d], batches[(size_t)d], threadsPerDev); });
                for (auto &x : th) x.join();
            }
            double w0 = nowSec();
            for (int d = 0; d < used; d++) {
                for (auto &f : batches[(size_t)d].fibers) fwrite(f->rc.out.data(), 1, f->rc.out.size(), out);
                nReads += batches[(size_t)d].reads.size();
            }
            tWrite += nowSec() - w0;
        }
        fflush(out);
        tAlign = nowSec() - tAlign;
        if (A.verbose || A.passes > 1 || getenv("YAHA_B200_STATS")) {
            ya_counters c{};
            double seed = 0, dp = 0, host = 0, upl = 0; uint64_t jobs = 0, rounds = 0, cells = 0, launches = 0, probes = 0, hits = 0, fragsAll = 0;
            double msdp = 0, msseed = 0, mstb = 0;
            for (Device &d : devs) {
                ya_get_counters(d.ctx, &c);
                seed += d.tSeed; dp += d.tDp; host += d.tHost; upl += d.tUpload; jobs += d.nJobs; rounds += d.nRounds; cells += c.dp_cells;
                msdp += c.ms_dp; msseed += c.ms_seed; mstb += c.ms_traceback; launches += c.launches; probes += c.probes; hits += c.hits;
                fragsAll += c.frags_all;
            }
            fprintf(stderr, "{\"pass\": %d, \"reads\": %llu, \"align_s\": %.5f, \"open_s\": %.3f, \"reads_per_s\": %.1f, \"read_parse_s\": %.5f, "
                    "\"upload_s\": %.5f, \"write_s\": %.5f, \"seed_wall_s\": %.5f, \"dp_wall_s\": %.5f, \"host_wall_s\": %.5f, \"dp_jobs\": %llu, "
                    "\"dp_rounds\": %llu, \"dp_cells\": %llu, \"dev_ms_seed\": %.3f, \"dev_ms_dp\": %.3f, \"dev_ms_traceback\": %.3f, "
                    "\"launches\": %llu, \"probes\": %llu, \"hits\": %llu, \"frags_all\": %llu, \"gpus\": %d, \"threads\": %d}\n",
                    pass, (unsigned long long)nReads, tAlign, tOpen, nReads / std::max(tAlign, 1e-9), tRead, upl, tWrite, seed, dp, host,
                    (unsigned long long)jobs, (unsigned long long)rounds, (unsigned long long)cells, msseed, msdp, mstb,
                    (unsigned long long)launches, (unsigned long long)probes, (unsigned long long)hits, (unsigned long long)fragsAll, nDev, nThreads);
        }
    }
    if (out != stdout) fclose(out); else fflush(out);
    qr.close();
    for (Device &d : devs) ya_close(d.ctx);
    return 0;
}

int runIndex(const Args &A)
{
    (void)A;
    fprintf(stderr, "yaha_b200: index creation (-g) is provided by the Python front end (yaha_b200.refio / Aligner(index=None));\n"
                    "the alignment hot path reads the reference's .nib2 and index files unchanged.\n");
    return 2;
}

}  // namespace yh
