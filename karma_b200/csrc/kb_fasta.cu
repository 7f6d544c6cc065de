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
#include <vector>

struct kb_fasta {
    std::vector<uint8_t> img;      // file image
    int64_t n_records, total_bases, total_key_bytes;
};

namespace {

struct Line { int64_t beg, end; };   // [beg, end) without the terminator

// next line starting at p; returns false at end of image
inline bool next_line(const std::vector<uint8_t>& img, int64_t& p, Line& ln) {
    const int64_t n = (int64_t)img.size();
    if (p >= n) return false;
    int64_t e = p;
    while (e < n && img[e] != '\n' && img[e] != '\r') ++e;
    ln.beg = p; ln.end = e;
    if (e < n) { if (img[e] == '\r' && e + 1 < n && img[e + 1] == '\n') e += 2; else e += 1; }
    p = e;
    return true;
}

inline int64_t key_end(const std::vector<uint8_t>& img, const Line& ln) {
    int64_t e = ln.beg;
    while (e < ln.end && img[e] != ' ') ++e;
    return e;
}

// One walk over the records.  With null outputs it only counts.
int walk(kb_fasta* f, uint8_t* bases, int64_t* offsets, int32_t* key_len, uint8_t* keys, int64_t* key_offsets) {
    const std::vector<uint8_t>& img = f->img;
    int64_t p = 0, rec = 0, nb = 0, nkb = 0;
    Line ln;
    // first line is the first header whatever it looks like (an empty file gives the key "")
    Line first{0, 0};
    next_line(img, p, first);
    auto open_record = [&](const Line& h) {
        const int64_t ke = key_end(img, h);
        if (keys) memcpy(keys + nkb, img.data() + h.beg, (size_t)(ke - h.beg));
        if (key_offsets) key_offsets[rec] = nkb;
        if (key_len) key_len[rec] = (int32_t)(ke - h.beg);
        if (offsets) offsets[rec] = nb;
        nkb += ke - h.beg;
    };
    open_record(first);
    while (next_line(img, p, ln)) {
        if (ln.end > ln.beg && img[ln.beg] == '>') {
            ++rec;
            open_record(ln);
        } else {
            if (bases) memcpy(bases + nb, img.data() + ln.beg, (size_t)(ln.end - ln.beg));
            nb += ln.end - ln.beg;
        }
    }
    ++rec;
    if (offsets) offsets[rec] = nb;
    if (key_offsets) key_offsets[rec] = nkb;
    f->n_records = rec; f->total_bases = nb; f->total_key_bytes = nkb;
    return KB_OK;
}

}  // namespace

extern "C" int kb_fasta_open(const char* path, kb_fasta** out, int64_t* n_records, int64_t* total_bases,
                             int64_t* total_key_bytes) {
    KB_CHECK_ARG(path && out, "null pointer");
    *out = nullptr;
    FILE* fp = fopen(path, "rb");
    if (!fp) { kb_set_error("cannot open %s", path); return KB_EINVAL; }
    kb_fasta* f = new kb_fasta();
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    f->img.resize(sz > 0 ? (size_t)sz : 0);
    if (sz > 0 && fread(f->img.data(), 1, (size_t)sz, fp) != (size_t)sz) {
        fclose(fp); delete f;
        kb_set_error("short read on %s", path);
        return KB_EINVAL;
    }
    fclose(fp);
    for (uint8_t b : f->img)
        if (b >= 0x80) { delete f; kb_set_error("%s holds non-ASCII bytes", path); return KB_EUNSUPPORTED; }
    walk(f, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (n_records) *n_records = f->n_records;
    if (total_bases) *total_bases = f->total_bases;
    if (total_key_bytes) *total_key_bytes = f->total_key_bytes;
    *out = f;
    return KB_OK;
}

extern "C" int kb_fasta_fill(kb_fasta* f, uint8_t* h_bases, int64_t* h_offsets, int32_t* h_key_len,
                             uint8_t* h_keys, int64_t* h_key_offsets) {
    KB_CHECK_ARG(f && h_offsets && h_key_len, "null pointer");
    return walk(f, h_bases, h_offsets, h_key_len, h_keys, h_key_offsets);
}

extern "C" int kb_fasta_close(kb_fasta* f) {
    delete f;
    return KB_OK;
}
