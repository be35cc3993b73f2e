// formats.cpp -- readers for yaha's on-disk formats and the FASTA/FASTQ query reader.
// Follows: loadBaseSequences Compress.c:76-134, normalizeBaseSequences / maxROff BaseSeq.c:115-125,
//          findBaseSequenceNum BaseSeq.c:81-90, index header Query.c:594-626,
//          code tables Math.c:141-157, readNextQuery Query.c:102-228, openQueryFile Query.c:63-74.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "host.hpp"
#if defined(__x86_64__)
#include <tmmintrin.h>
#endif

namespace yh {

const char kCharOfCode[16] = {'T', 'C', 'A', 'G', 'N', 'B', 'D', 'H', 'K', 'M', 'R', 'S', 'V', 'W', 'X', 'Y'};
const char kCompCharOfCode[16] = {'A', 'G', 'T', 'C', 'N', 'V', 'H', 'D', 'M', 'K', 'Y', 'S', 'B', 'W', 'X', 'R'};
const uint8_t kCompCode[16] = {2, 3, 0, 1, 4, 12, 7, 6, 9, 8, 15, 11, 5, 13, 14, 10};

int codeOfChar(int c)
{
    static int8_t table[128];
    static bool ready = false;
    if (!ready) {
        memset(table, 14, sizeof table);
        for (int i = 0; i < 16; i++) { table[(int)kCharOfCode[i]] = (int8_t)i; table[(int)kCharOfCode[i] + 32] = (int8_t)i; }
        table['U'] = table['u'] = 0;
        ready = true;
    }
    return (c >= 0 && c < 128) ? table[c] : 14;
}

static const void *mapFile(const std::string &path, size_t &len, std::string &err)
{
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = (errno == ENOENT) ? "File '" + path + "' does not exist." : "cannot open '" + path + "'"; return nullptr; }
    struct stat st;
    if (fstat(fd, &st) != 0) { err = "cannot stat '" + path + "'"; close(fd); return nullptr; }
    len = (size_t)st.st_size;
    void *p = mmap(nullptr, len, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { err = "cannot mmap '" + path + "'"; return nullptr; }
    return p;
}

bool Genome::load(const std::string &path, std::string &err)
{
    map = mapFile(path, mapLen, err);
    if (!map) return false;
    const uint32_t *h = (const uint32_t *)map;
    const int ver = (int)h[1];
    if (h[0] != 0x01020304u || (ver != 1 && ver != 2)) { err = "Input nib2 file bad header format."; return false; }
    const int blk = ver == 2 ? 16 : 12;
    const int n = (int)h[3];
    bases = (const uint8_t *)map + h[2];
    nBaseBytes = mapLen - h[2];
    const char *names = (const char *)map + 16 + blk * n + 4;
    const uint32_t *rec = h + 4;
    seqs.clear();
    for (int i = 0; i < n; i++) {
        BaseSeq b;
        b.start = rec[0] * 2;                        // byte offset -> base offset (BaseSeq.c:115-119)
        b.length = rec[1];
        if (ver == 1) { uint32_t info = rec[2]; b.name.assign(names + (uint16_t)(info >> 16), info & 0xFFFF); rec += 3; }
        else { b.name.assign(names + rec[2], rec[3]); rec += 4; }
        seqs.push_back(b);
    }
    if (seqs.empty()) { err = "nib2 file without sequences"; return false; }
    maxROff = seqs.back().start + seqs.back().length;
    return true;
}

int Genome::findSeq(uint32_t off) const
{
    for (size_t i = 0; i < seqs.size(); i++)
        if (off >= seqs[i].start && off < seqs[i].start + seqs[i].length) return (int)i;
    return -1;
}

bool IndexFile::load(const std::string &path, std::string &err)
{
    map = mapFile(path, mapLen, err);
    if (!map) return false;
    const uint32_t *u = (const uint32_t *)map;
    if (mapLen < 16 || (int)u[0] != -1) { err = "Index file version is out of date.\nPlease remake index file and try again."; return false; }
    wordLen = (int)u[1]; maxHits = (int)u[2];
    if (wordLen < 1 || wordLen > 16) { err = "Index file has an unsupported word length."; return false; }
    nSo = ((size_t)1 << (2 * wordLen)) + 1;
    so = u + 4;
    roa = so + nSo;
    if (mapLen < 16 + nSo * 4) { err = "Index file is truncated."; return false; }
    nRoa = (mapLen - 16 - nSo * 4) / 4;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Query reader.  Same record semantics as readNextQuery (Query.c:102-228) -- id cut at newline, spaces
// -> '_', only '\n' skipped inside sequences, the break character ('>' / '+') ends a sequence wherever
// it occurs, over-long and too-short reads skipped with the reference's warnings -- but reading
// through a private 4 MiB buffer with whole-line copies instead of one getc per character.
struct QueryReader::Buf {
    std::vector<char> b;
    size_t p = 0, e = 0;
    bool eof = false;
    Buf() : b((size_t)4 << 20) {}
    bool fill(FILE *f)
    {
        if (eof) return false;
        p = 0;
        e = fread(b.data(), 1, b.size(), f);
        if (e == 0) { eof = true; return false; }
        return true;
    }
    inline int get(FILE *f)
    {
        if (p == e && !fill(f)) return EOF;
        return (unsigned char)b[p++];
    }
};

bool QueryReader::open(const std::string &path, std::string &err)
{
    // "stdin" / "-": the query stream is standard input (Main.c:173-178, Query.c:63-74)
    f = (path == "stdin" || path == "-") ? stdin : fopen(path.c_str(), "r");
    if (!f) { err = "Failure to open input file: " + path + ".  Error number:" + std::to_string(errno); return false; }
    setvbuf(f, nullptr, _IONBF, 0);
    buf = new Buf();
    fastq = (buf->get(f) == '@');                    // also positions the stream after the first marker
    return true;
}

void QueryReader::close()
{
    if (f && f != stdin) fclose(f);
    f = nullptr;
    delete buf;
    buf = nullptr;
}

static void readToChar(QueryReader::Buf &B, FILE *f, int target, bool afterNewline)          // Query.c:52-61
{
    int prev = 0;
    for (;;) {
        int c = B.get(f);
        if ((c == target && (!afterNewline || prev == '\n')) || c == EOF) return;
        prev = c;
    }
}

bool QueryReader::next(Read &r)
{
    Buf &B = *buf;
    for (;;) {
        r.id.clear(); r.fwd.clear(); r.qual.clear();
        int idChars = 0;
        for (;;) {                                   // id line
            if (B.p == B.e && !B.fill(f)) break;
            const char *s = B.b.data() + B.p;
            const char *nl = (const char *)memchr(s, '\n', B.e - B.p);
            size_t len = nl ? (size_t)(nl - s) : B.e - B.p;
            for (size_t k = 0; k < len; k++, idChars++)
                if (idChars < 200) r.id.push_back(s[k] == ' ' ? '_' : s[k]);
            B.p += len;
            if (nl) { B.p++; break; }
        }
        if (idChars > 200)
            fprintf(stderr, "Warning, Query Id length of %d exceeds maximum length %d.  Id will be truncated.\n", idChars, 200);
        const int brk = fastq ? '+' : '>';
        bool fail = false;
        for (;;) {                                   // sequence lines
            if (B.p == B.e && !B.fill(f)) break;
            const char *s = B.b.data() + B.p;
            size_t avail = B.e - B.p;
            const char *nl = (const char *)memchr(s, '\n', avail);
            size_t len = nl ? (size_t)(nl - s) : avail;
            const char *bk = (const char *)memchr(s, brk, len);
            if (bk) len = (size_t)(bk - s);
            if ((int)(r.fwd.size() + len) > maxLen) {
                size_t room = (size_t)maxLen - r.fwd.size();
                r.fwd.append(s, room);
                B.p += room;
                fprintf(stderr, "Warning.  Query sequence exceeds maximum length of %d.  Query will be skipped.\n", maxLen);
                B.p += 1;                            // the character that did not fit is consumed (Query.c:144-155)
                readToChar(B, f, brk, false);
                fail = true;
                break;
            }
            r.fwd.append(s, len);
            B.p += len;
            if (bk) { B.p++; break; }
            if (nl) B.p++;
        }
        if (fastq) {
            readToChar(B, f, '\n', false);
            int prev = 0;
            for (;;) {
                int c = B.get(f);
                if ((c == '@' && prev == '\n') || c == EOF) break;
                prev = c;
                if (c == '\n') continue;
                if ((int)r.qual.size() >= maxLen) {
                    fprintf(stderr, "Warning.  Quality score sequence exceeds maximum length of %d.  Query will be skipped.\n", maxLen);
                    readToChar(B, f, '@', true);
                    fail = true;
                    break;
                }
                r.qual.push_back((char)c);
            }
            if (r.fwd.size() != r.qual.size()) {
                fprintf(stderr, "Warning.  Query sequence (%d) and quality score sequence (%d) have different lengths in fastq file."
                        "  Query will be skipped.\n", (int)r.fwd.size(), (int)r.qual.size());
                fail = true;
            }
        }
        const int n = (int)r.fwd.size();
        if (n > 0 && n < wordLen) {
            fprintf(stderr, "Query length must be at least wordlen bases long. Query will be skipped.\n");
            fail = true;
        }
        if (fail) continue;
        // A record without bases ends the input like EOF does (the caller's loop runs while queryLen > 0, Query.c:306) -- except
        // when it is the first record readNextQuery returns: the reference then reads once more ("make sure we have read a
        // first query", Query.c:304, after main's own first read at :637).
        const bool firstReturn = !returnedOnce;
        returnedOnce = true;
        if (n == 0) { if (firstReturn) continue; return false; }
        // codes are derived later, off the reader thread: forward by the pipeline that uploads the batch
        // (Read::encode), reverse-complement by the worker that owns the read (Read::finish)
        r.fcode.clear();
        r.rcode.clear(); r.rev.clear();
        return true;
    }
}

bool RecordSlicer::open(const std::string &path, std::string &err)
{
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "Failure to open input file: " + path + ".  Error number:" + std::to_string(errno); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0) { err = "Failure to stat input file: " + path; ::close(fd); return false; }
    len = (size_t)st.st_size; pos = 0; done = true; base = nullptr;
    if (len > 0) {
        // small inputs are mapped with their pages in place; large ones are read ahead as the slicer walks through them
        void *m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE | (len <= ((size_t)256 << 20) ? MAP_POPULATE : 0), fd, 0);
        if (m == MAP_FAILED) { err = "Failure to map input file: " + path; ::close(fd); return false; }
        madvise(m, len, MADV_SEQUENTIAL);
        base = (const char *)m;
        pos = 1;                                     // the first character is taken as the first marker (QueryReader::open)
        done = false;
    }
    ::close(fd);
    return true;
}

void RecordSlicer::close()
{
    if (base) munmap((void *)base, len);
    base = nullptr; len = pos = 0; done = true;
}

bool RecordSlicer::next(const char *&s, size_t &n)
{
  for (;;) {
    if (done) return false;
    if (pos > len) { done = true; return false; }
    s = base + pos;
    // the id line runs to its newline whatever it contains (Query.c:111-135); only the sequence lines
    // behind it end at the next marker (Query.c:137-158)
    const char *nl = (const char *)memchr(s, '\n', len - pos);
    const size_t idEnd = nl ? (size_t)(nl - s) + 1 : len - pos;
    const char *bk = (const char *)memchr(s + idEnd, '>', len - pos - idEnd);
    n = bk ? (size_t)(bk - s) : len - pos;
    pos += n + 1;
    // a record without sequence characters ends the input (Query.c:222): id line, then nothing but newlines
    size_t q = idEnd;
    while (q < n && s[q] == '\n') q++;
    if (q >= n) {
        // ... except when no record before it was a good one (the parser skips records that are too long or shorter than
        // wordLen): it is then the first record readNextQuery RETURNS, and the reference reads once more (Query.c:304,637).
        if (!anyGood && !skippedEmpty) { skippedEmpty = true; continue; }
        done = true; return false;
    }
    if (!anyGood && !skippedEmpty) {
        size_t m = 0;
        for (size_t k = idEnd; k < n; k++) m += s[k] != '\n';
        if ((int64_t)m >= wordLen && (int64_t)m <= maxLen) anyGood = true;
    }
    return true;
  }
}

int parseFastaRecord(const char *s, size_t n, Read &r, int maxLen, int wordLen)
{
    r.id.clear(); r.fwd.clear(); r.qual.clear(); r.fcode.clear(); r.rcode.clear(); r.rev.clear();
    const char *nl = (const char *)memchr(s, '\n', n);
    const size_t idLen = nl ? (size_t)(nl - s) : n;
    const size_t keep = idLen < 200 ? idLen : 200;
    r.id.assign(s, keep);
    for (size_t k = 0; k < keep; k++) if (r.id[k] == ' ') r.id[k] = '_';
    if (idLen > 200)
        fprintf(stderr, "Warning, Query Id length of %d exceeds maximum length %d.  Id will be truncated.\n", (int)idLen, 200);
    size_t p = nl ? idLen + 1 : n;
    bool fail = false;
    while (p < n) {
        const char *nl2 = (const char *)memchr(s + p, '\n', n - p);
        const size_t l = nl2 ? (size_t)(nl2 - (s + p)) : n - p;
        if ((int)(r.fwd.size() + l) > maxLen) {
            r.fwd.append(s + p, (size_t)maxLen - r.fwd.size());
            fprintf(stderr, "Warning.  Query sequence exceeds maximum length of %d.  Query will be skipped.\n", maxLen);
            fail = true;
            break;
        }
        r.fwd.append(s + p, l);
        p += l;
        if (nl2) p++;
    }
    const int m = (int)r.fwd.size();
    if (m > 0 && m < wordLen) {
        fprintf(stderr, "Query length must be at least wordlen bases long. Query will be skipped.\n");
        fail = true;
    }
    return (fail || m == 0) ? 0 : 1;
}

// The same record straight into flat buffers (the page-locked inputs of ya_align_batch): id (cut to 200 characters, blanks
// replaced) appended at ids + *idLen, sequence characters at chars + *seqLen.  Returns 1 and advances both lengths, or 0 (record
// skipped with the reference's warning: too long / shorter than wordLen / empty) and leaves them alone.  chars must have room
// for the record's n bytes, ids for 200.
int parseFastaRecordInto(const char *s, size_t n, char *chars, size_t *seqLen, char *ids, size_t *idLen, int maxLen, int wordLen)
{
    const char *nl = (const char *)memchr(s, '\n', n);
    const size_t idl = nl ? (size_t)(nl - s) : n;
    const size_t keep = idl < 200 ? idl : 200;
    char *idDst = ids + *idLen;
    for (size_t k = 0; k < keep; k++) idDst[k] = s[k] == ' ' ? '_' : s[k];
    if (idl > 200)
        fprintf(stderr, "Warning, Query Id length of %d exceeds maximum length %d.  Id will be truncated.\n", (int)idl, 200);
    char *dst = chars + *seqLen;
    size_t m = 0, p = nl ? idl + 1 : n;
    bool fail = false;
    while (p < n) {
        const char *nl2 = (const char *)memchr(s + p, '\n', n - p);
        const size_t l = nl2 ? (size_t)(nl2 - (s + p)) : n - p;
        if ((int)(m + l) > maxLen) {
            fprintf(stderr, "Warning.  Query sequence exceeds maximum length of %d.  Query will be skipped.\n", maxLen);
            fail = true;
            break;
        }
        memcpy(dst + m, s + p, l);
        m += l;
        p += l;
        if (nl2) p++;
    }
    if (!fail && m > 0 && (int)m < wordLen) {
        fprintf(stderr, "Query length must be at least wordlen bases long. Query will be skipped.\n");
        fail = true;
    }
    if (fail || m == 0) return 0;
    *seqLen += m; *idLen += keep;
    return 1;
}

// codeOfChar for 16 characters per step.  A letter's code depends on its low five bits only (upper and lower case agree), so
// the 256-entry table folds into two 16-entry byte shuffles; everything that is not an ASCII letter is X (14) like the table
// says.  The routine is compared with the table on all 256 byte values before it is first used, and left unused otherwise.
#if defined(__x86_64__)
__attribute__((target("ssse3"))) static size_t encode16(const unsigned char *src, uint8_t *dst, size_t n, const uint8_t *lut32)
{
    const __m128i lutLo = _mm_loadu_si128((const __m128i *)lut32), lutHi = _mm_loadu_si128((const __m128i *)(lut32 + 16));
    const __m128i c20 = _mm_set1_epi8(0x20), ca = _mm_set1_epi8('a'), c25 = _mm_set1_epi8(25), c15 = _mm_set1_epi8(15);
    const __m128i c31 = _mm_set1_epi8(31), cX = _mm_set1_epi8(14);
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128((const __m128i *)(src + i));
        const __m128i t = _mm_sub_epi8(_mm_or_si128(v, c20), ca);
        const __m128i letter = _mm_cmpeq_epi8(_mm_min_epu8(t, c25), t);            // (v | 0x20) - 'a' < 26, unsigned
        const __m128i idx = _mm_and_si128(v, c31);
        const __m128i high = _mm_cmpgt_epi8(idx, c15);
        const __m128i low4 = _mm_and_si128(idx, c15);
        const __m128i code = _mm_or_si128(_mm_andnot_si128(high, _mm_shuffle_epi8(lutLo, low4)), _mm_and_si128(high, _mm_shuffle_epi8(lutHi, low4)));
        _mm_storeu_si128((__m128i *)(dst + i), _mm_or_si128(_mm_and_si128(letter, code), _mm_andnot_si128(letter, cX)));
    }
    return i;
}
#endif

void Read::encode()                                 // Query.c:161-163 (codeOfChar per base)
{
    static uint8_t codeTab[256];
    static const bool tabReady = [] { for (int c = 0; c < 256; c++) codeTab[c] = (uint8_t)codeOfChar(c < 128 ? c : 0); return true; }();
    (void)tabReady;
    const size_t n = fwd.size();
    if (fcode.size() == n) return;
    fcode.resize(n);
    const unsigned char *src = (const unsigned char *)fwd.data();
    uint8_t *fc = fcode.data();
    size_t i = 0;
#if defined(__x86_64__)
    static uint8_t lut32[32];
    static const bool wide = [] {
        if (!__builtin_cpu_supports("ssse3")) return false;
        for (int k = 0; k < 32; k++) lut32[k] = codeTab[k >= 1 && k <= 26 ? 'A' + k - 1 : 0];
        unsigned char all[256]; uint8_t got[256];
        for (int c = 0; c < 256; c++) all[c] = (unsigned char)c;
        if (encode16(all, got, 256, lut32) != 256) return false;
        for (int c = 0; c < 256; c++) if (got[c] != codeTab[c]) return false;
        return true;
    }();
    if (wide) i = encode16(src, fc, n, lut32);
#endif
    for (; i < n; i++) fc[i] = codeTab[src[i]];
}

// reverse complement 16 bases per step: both tables (complement of a 4-bit code, character of a code) have 16 entries,
// i.e. each is one byte shuffle; a third shuffle reverses the block
#if defined(__x86_64__)
__attribute__((target("ssse3"))) static int revcomp16(const uint8_t *f, uint8_t *rc, char *rv, int n)
{
    const __m128i comp = _mm_loadu_si128((const __m128i *)kCompCode);
    const __m128i chr = _mm_loadu_si128((const __m128i *)kCharOfCode);
    const __m128i flip = _mm_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(f + i)), flip);
        const __m128i c = _mm_shuffle_epi8(comp, v);
        _mm_storeu_si128((__m128i *)(rc + n - 16 - i), c);
        _mm_storeu_si128((__m128i *)(rv + n - 16 - i), _mm_shuffle_epi8(chr, c));
    }
    return i;
}
#endif

void Read::finish()                                 // Query.c:164-167
{
    const int n = (int)fcode.size();
    if ((int)rcode.size() == n) return;
    rcode.resize((size_t)n); rev.resize((size_t)n);
    int i = 0;
#if defined(__x86_64__)
    static const bool ssse3 = __builtin_cpu_supports("ssse3");
    if (ssse3) i = revcomp16(fcode.data(), rcode.data(), &rev[0], n);
#endif
    for (; i < n; i++) {
        const uint8_t cc = kCompCode[fcode[(size_t)i]];
        rcode[(size_t)(n - 1 - i)] = cc;
        rev[(size_t)(n - 1 - i)] = kCharOfCode[cc];
    }
}

}  // namespace yh
