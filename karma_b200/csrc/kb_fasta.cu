// kb_fasta.cu -- native FASTA reader/packer (host code only).
//
// Replaces read_fasta_file (/root/reference/karma/karma.py:40-61) plus the Python
// dict -> buffer marshalling in front of the GPU path: one pass over the file image emits
// the packed form the C ABI consumes (bases back to back, offsets, key lengths) together
// with the header keys, so that the Python layer can still hand karma.py the same
// OrderedDict{">name": sequence}.
//
// Semantics restated from karma.py:40-61 (Python text mode => universal newlines):
//   * lines end at "\n", "\r\n" or "\r";
//   * the FIRST line is always a header: key = line.rstrip("\n").split(" ")[0]
//     (the leading '>' is kept; only ' ' splits, tabs do not);
//   * every later line starting with '>' closes the current record and opens a new one;
//     any other line is appended to the sequence with its line ending removed;
//   * the last record is always emitted (an empty file yields one record "" -> "").
// Duplicate keys are the caller's business (a Python dict keeps the first position and
// the last value); bytes >= 0x80 are rejected (the reference decodes UTF-8, so byte and
// character counts would differ).
#include "kb_common.cuh"
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

// The file image is cut into chunks that start at line starts; every chunk is scanned by its own
// thread (memchr line scanning).  Pass 1 (open) lists the records each chunk opens and the bases
// it adds to the record that is open when the chunk starts; a short serial prefix gives every
// chunk its first record number and its base / key offsets; pass 2 (fill) copies in parallel.
struct kb_fasta {
    uint8_t* img = nullptr;        // file image (malloc)
    int64_t size = 0;
    int64_t n_records = 0, total_bases = 0, total_key_bytes = 0;
    struct Chunk {
        int64_t beg = 0, end = 0;  // [beg, end), beg is a line start
        int64_t headers = 0;       // records opened inside the chunk
        int64_t bases = 0;         // bases of all sequence lines of the chunk
        int64_t key_bytes = 0;
        int64_t rec0 = 0, base0 = 0, key0 = 0;   // prefix: record open at chunk start, offsets
    };
    std::vector<Chunk> chunks;
    ~kb_fasta() { free(img); }
};

namespace {

struct Line { int64_t beg, end; };   // [beg, end) without the terminator

// next line starting at p (p < lim); lines end at "\n", "\r\n" or "\r"
inline void next_line(const uint8_t* img, int64_t n, int64_t& p, Line& ln) {
    const uint8_t* s = img + p;
    const uint8_t* nl = static_cast<const uint8_t*>(memchr(s, '\n', (size_t)(n - p)));
    const int64_t e_nl = nl ? (int64_t)(nl - img) : n;
    // a '\r' before that '\n' ends the line earlier (rare: one more memchr over the same bytes)
    const uint8_t* cr = static_cast<const uint8_t*>(memchr(s, '\r', (size_t)(e_nl - p)));
    int64_t e = cr ? (int64_t)(cr - img) : e_nl;
    ln.beg = p; ln.end = e;
    if (e < n) { if (img[e] == '\r' && e + 1 < n && img[e + 1] == '\n') e += 2; else e += 1; }
    p = e;
}

inline int64_t key_end(const uint8_t* img, const Line& ln) {
    const uint8_t* sp = static_cast<const uint8_t*>(memchr(img + ln.beg, ' ', (size_t)(ln.end - ln.beg)));
    return sp ? (int64_t)(sp - img) : ln.end;
}

inline bool is_line_start(const uint8_t* img, int64_t p) {
    return p == 0 || img[p - 1] == '\n' || (img[p - 1] == '\r' && img[p] != '\n');
}

// One chunk.  With null outputs it only counts (pass 1); otherwise it writes the records it opens
// and the bases of its sequence lines (pass 2).  The very first line of the file is a header
// whatever it looks like.
void walk_chunk(const kb_fasta* f, kb_fasta::Chunk& c, bool count_only, uint8_t* bases, int64_t* offsets, int32_t* key_len,
                uint8_t* keys, int64_t* key_offsets) {
    const uint8_t* img = f->img;
    int64_t p = c.beg, rec = c.rec0, nb = c.base0, nkb = c.key0;
    int64_t headers = 0, nbases = 0, nkeys = 0;
    Line ln;
    while (p < c.end) {
        const bool first_of_file = p == 0;
        next_line(img, f->size, p, ln);
        if (first_of_file || (ln.end > ln.beg && img[ln.beg] == '>')) {
            const int64_t kl = key_end(img, ln) - ln.beg;
            ++headers;
            if (!count_only) {
                const int64_t r = rec + headers;          // record numbers are 1-based here: rec0 counts the headers before
                if (keys) memcpy(keys + nkb + nkeys, img + ln.beg, (size_t)kl);
                if (key_offsets) key_offsets[r - 1] = nkb + nkeys;
                key_len[r - 1] = (int32_t)kl;
                offsets[r - 1] = nb + nbases;
            }
            nkeys += kl;
        } else {
            if (!count_only && bases) memcpy(bases + nb + nbases, img + ln.beg, (size_t)(ln.end - ln.beg));
            nbases += ln.end - ln.beg;
        }
    }
    if (count_only) { c.headers = headers; c.bases = nbases; c.key_bytes = nkeys; }
}

template <class F>
void for_chunks(kb_fasta* f, F fn) {
    const size_t n = f->chunks.size();
    if (n <= 1) { for (auto& c : f->chunks) fn(c); return; }
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t i = 0; i < n; ++i) th.emplace_back([&, i] { fn(f->chunks[i]); });
    for (auto& t : th) t.join();
}

bool has_high_bytes(const uint8_t* p, int64_t n) {
    uint64_t acc = 0;
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); acc |= w; }
    for (; i < n; ++i) acc |= (uint64_t)p[i] << 0;
    return (acc & 0x8080808080808080ull) != 0;
}

}  // namespace

int64_t kb_host_threads() {
    int64_t want = (int64_t)std::thread::hardware_concurrency();
    if (const char* e = getenv("KB_HOST_THREADS")) want = atoll(e);
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    return want;
}

// Whole file into a malloc'ed image (caller frees).  Files of 64 MB and more are read by all host
// threads at once: the page-cache copy and the first-touch faults are the cost, not the disk.
int kb_host_read_file(const char* path, uint8_t** img, int64_t* size) {
    *img = nullptr; *size = 0;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { kb_set_error("cannot open %s", path); return KB_EINVAL; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); kb_set_error("%s is not a regular file", path); return KB_EINVAL; }
    const int64_t sz = (int64_t)sb.st_size;
    uint8_t* buf = static_cast<uint8_t*>(malloc((size_t)sz + 1));
    if (!buf) { close(fd); kb_set_error("out of host memory reading %s", path); return KB_EINVAL; }
    const int64_t readers = sz >= (64LL << 20) ? kb_host_threads() : 1;
    std::vector<int> ok((size_t)readers, 1);
    auto slice = [&](int64_t t) {
        int64_t at = sz * t / readers;
        const int64_t end = sz * (t + 1) / readers;
        while (at < end) {
            const ssize_t got = pread(fd, buf + at, (size_t)(end - at), (off_t)at);
            if (got <= 0) { ok[(size_t)t] = 0; return; }
            at += got;
        }
    };
    if (readers == 1) slice(0);
    else {
        std::vector<std::thread> th;
        for (int64_t t = 0; t < readers; ++t) th.emplace_back(slice, t);
        for (auto& t : th) t.join();
    }
    close(fd);
    for (int o : ok)
        if (!o) { free(buf); kb_set_error("short read on %s", path); return KB_EINVAL; }
    *img = buf; *size = sz;
    return KB_OK;
}

extern "C" int kb_fasta_open(const char* path, kb_fasta** out, int64_t* n_records, int64_t* total_bases,
                             int64_t* total_key_bytes) {
    KB_CHECK_ARG(path && out, "null pointer");
    *out = nullptr;
    kb_fasta* f = new kb_fasta();
    { const int rc = kb_host_read_file(path, &f->img, &f->size); if (rc) { delete f; return rc; } }
    int64_t want = kb_host_threads();
    // chunks of >= 4 MB starting at line starts, at most one per hardware thread (KB_HOST_THREADS overrides)
    const int64_t min_chunk = getenv("KB_FASTA_MIN_CHUNK") ? atoll(getenv("KB_FASTA_MIN_CHUNK")) : (4LL << 20);
    if (want > f->size / (min_chunk > 0 ? min_chunk : 1) + 1) want = f->size / (min_chunk > 0 ? min_chunk : 1) + 1;
    int64_t prev = 0;
    for (int64_t t = 1; t <= want; ++t) {
        int64_t cut = t == want ? f->size : f->size * t / want;
        while (cut < f->size && !is_line_start(f->img, cut)) ++cut;
        if (cut > prev || t == want) {
            kb_fasta::Chunk c; c.beg = prev; c.end = cut;
            if (c.end > c.beg || f->chunks.empty()) f->chunks.push_back(c);
            prev = cut;
        }
    }
    // non-ASCII check and pass 1, chunk-parallel
    std::vector<uint8_t> high(f->chunks.size(), 0);
    {
        const size_t n = f->chunks.size();
        std::vector<std::thread> th;
        auto work = [&](size_t i) {
            kb_fasta::Chunk& c = f->chunks[i];
            high[i] = has_high_bytes(f->img + c.beg, c.end - c.beg);
            if (!high[i]) walk_chunk(f, c, true, nullptr, nullptr, nullptr, nullptr, nullptr);
        };
        if (n <= 1) { for (size_t i = 0; i < n; ++i) work(i); }
        else { for (size_t i = 0; i < n; ++i) th.emplace_back(work, i); for (auto& t : th) t.join(); }
    }
    for (uint8_t h : high)
        if (h) { delete f; kb_set_error("%s holds non-ASCII bytes", path); return KB_EUNSUPPORTED; }
    int64_t rec = 0, nb = 0, nkb = 0;
    for (auto& c : f->chunks) {
        c.rec0 = rec; c.base0 = nb; c.key0 = nkb;
        rec += c.headers; nb += c.bases; nkb += c.key_bytes;
    }
    if (f->size == 0) rec = 1;                                 // an empty file yields one record "" -> ""
    f->n_records = rec; f->total_bases = nb; f->total_key_bytes = nkb;
    if (n_records) *n_records = f->n_records;
    if (total_bases) *total_bases = f->total_bases;
    if (total_key_bytes) *total_key_bytes = f->total_key_bytes;
    *out = f;
    return KB_OK;
}

extern "C" int kb_fasta_fill(kb_fasta* f, uint8_t* h_bases, int64_t* h_offsets, int32_t* h_key_len,
                             uint8_t* h_keys, int64_t* h_key_offsets) {
    KB_CHECK_ARG(f && h_offsets && h_key_len, "null pointer");
    if (f->size == 0) {                                        // the empty record
        h_offsets[0] = 0; h_offsets[1] = 0; h_key_len[0] = 0;
        if (h_key_offsets) { h_key_offsets[0] = 0; h_key_offsets[1] = 0; }
        return KB_OK;
    }
    for_chunks(f, [&](kb_fasta::Chunk& c) { walk_chunk(f, c, false, h_bases, h_offsets, h_key_len, h_keys, h_key_offsets); });
    h_offsets[f->n_records] = f->total_bases;
    if (h_key_offsets) h_key_offsets[f->n_records] = f->total_key_bytes;
    return KB_OK;
}

extern "C" int kb_fasta_close(kb_fasta* f) {
    delete f;
    return KB_OK;
}
