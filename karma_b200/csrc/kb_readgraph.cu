// kb_readgraph.cu -- read-graph edge accumulation from salmon equivalence classes
// (a separate path from the k-mer front end).
//
// Replaces the loops of ReadGraph.from_equivalence_classes
// (/root/reference/karma/read_graph.py:61-148): per-contig read totals (:86-93), shared-read
// counts of every contig pair that co-occurs in an equivalence class (:99-115, a Python loop
// over itertools.combinations with has_edge/get_edge_data/add_edge per OCCURRENCE) and the
// normalised weight ((shared/totA)+(shared/totB))/2 (:117-131).
//
// GPU formulation (sort / reduce-by-key, HBM-bound integer work):
//   1. totals[id] += count               one thread per class, 64-bit global reductions
//   2. every pair occurrence -> key (min_id << 32 | max_id), payload = its sequence number in
//      the reference's iteration order; stable radix sort by key (CUB), so the first element
//      of every run is the pair's FIRST occurrence
//   3. reduce runs: shared = sum of counts, first = min sequence number
//   4. networkx yields graph.edges() grouped by the endpoint that was inserted first (here:
//      the smaller contig index), neighbours in first-occurrence order: a second sort by
//      (min_id, first) reproduces the reference's edge order exactly; zero-weight edges drop out
//   5. weight in IEEE fp64 with the reference's operation order (two divisions, add, halve)
// The host part parses eq_classes.txt (read_graph.py:75-82) into CSR arrays.
#include "kb_common.cuh"
#include <cub/cub.cuh>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct kb_eq {
    std::vector<uint8_t> names;            // contig names back to back
    std::vector<int64_t> name_off;         // n+1
    std::vector<int64_t> class_off;        // C+1
    std::vector<int32_t> ids;
    std::vector<int64_t> counts;           // C
    std::vector<uint8_t> skip;             // C: first token == "1" (read_graph.py:102)
};

namespace {

// temporary from the context's stream-ordered pool, returned to it (not to the driver) when it goes out of scope
struct DevBuf {
    static thread_local cudaMemPool_t pool;
    static thread_local cudaStream_t stream;
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, stream); }
    cudaError_t alloc(size_t bytes) { return cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, pool, stream); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
thread_local cudaMemPool_t DevBuf::pool = nullptr;
thread_local cudaStream_t DevBuf::stream = 0;

__global__ void rg_totals(const int64_t* __restrict__ class_off, const int32_t* __restrict__ ids,
                          const int64_t* __restrict__ counts, int64_t n_classes, unsigned long long* __restrict__ totals) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_classes) return;
    const unsigned long long cnt = (unsigned long long)counts[c];
    if (cnt == 0) return;
    for (int64_t t = class_off[c]; t < class_off[c + 1]; ++t) atomicAdd(&totals[ids[t]], cnt);
}

__global__ void rg_pair_counts(const int64_t* __restrict__ class_off, const uint8_t* __restrict__ skip, int64_t n_classes,
                               unsigned long long* __restrict__ n_pairs) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_classes) return;
    const unsigned long long s = (unsigned long long)(class_off[c + 1] - class_off[c]);
    n_pairs[c] = skip[c] ? 0ull : s * (s - 1) / 2;
}

// one warp per class: pair t of the class (itertools.combinations order) -> slot pair_off[c] + t
__global__ void rg_emit(const int64_t* __restrict__ class_off, const int32_t* __restrict__ ids,
                        const uint8_t* __restrict__ skip, const unsigned long long* __restrict__ pair_off,
                        int64_t n_classes, uint64_t* __restrict__ keys, uint32_t* __restrict__ seq, uint32_t* __restrict__ cls) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < n_classes; c += nwarps) {
        if (skip[c]) continue;
        const int64_t b = class_off[c];
        const int s = (int)(class_off[c + 1] - b);
        unsigned long long slot0 = pair_off[c];
        for (int i = 0; i < s - 1; ++i) {                       // row i of the combinations triangle: (i, i+1..s-1)
            const uint32_t a = (uint32_t)ids[b + i];
            for (int j = i + 1 + lane; j < s; j += 32) {
                const uint32_t bb = (uint32_t)ids[b + j];
                const unsigned long long slot = slot0 + (unsigned long long)(j - i - 1);
                const uint32_t lo = a < bb ? a : bb, hi = a < bb ? bb : a;
                keys[slot] = ((uint64_t)lo << 32) | hi;
                seq[slot] = (uint32_t)slot;
                cls[slot] = (uint32_t)c;
            }
            slot0 += (unsigned long long)(s - 1 - i);
        }
    }
}

__global__ void rg_gather(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, int64_t m, uint32_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) dst[i] = src[idx[i]];
}

// after the stable sort by key: run heads, per-element count
__global__ void rg_heads(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ cls, const int64_t* __restrict__ counts,
                         int64_t m, uint8_t* __restrict__ head, unsigned long long* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    head[i] = (i == 0) || keys[i] != keys[i - 1];
    cnt[i] = (unsigned long long)counts[cls[i]];
}

// second-order key (min_id, first occurrence) of every unique edge
__global__ void rg_order_keys(const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ first_seq, int64_t u,
                              uint64_t* __restrict__ okeys, uint32_t* __restrict__ oidx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    okeys[i] = (ukeys[i] & 0xffffffff00000000ull) | first_seq[i];
    oidx[i] = (uint32_t)i;
}

__global__ void rg_weights(const uint64_t* __restrict__ ukeys, const unsigned long long* __restrict__ shared,
                           const uint32_t* __restrict__ order, const unsigned long long* __restrict__ totals, int64_t u,
                           int32_t* __restrict__ ea, int32_t* __restrict__ eb, double* __restrict__ w,
                           unsigned long long* __restrict__ sh_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u) return;
    const uint32_t e = order[i];
    const uint64_t key = ukeys[e];
    const int32_t a = (int32_t)(key >> 32), b = (int32_t)(key & 0xffffffffu);
    const unsigned long long s = shared[e];
    ea[i] = a; eb[i] = b; sh_out[i] = s;
    // ((shared / totA) + (shared / totB)) / 2 -- read_graph.py:125-127, IEEE fp64, no contraction
    const double sd = (double)s;
    const double qa = s ? __ddiv_rn(sd, (double)totals[a]) : 0.0;
    const double qb = s ? __ddiv_rn(sd, (double)totals[b]) : 0.0;
    w[i] = __dmul_rn(__dadd_rn(qa, qb), 0.5);
}

bool parse_int(const uint8_t* p, const uint8_t* e, int64_t* out) {
    if (p == e) return false;
    bool neg = false;
    if (*p == '-') { neg = true; ++p; if (p == e) return false; }
    int64_t v = 0;
    for (; p < e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    *out = neg ? -v : v;
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host: eq_classes.txt parser (read_graph.py:75-82)
// ---------------------------------------------------------------------------------------
namespace {

// One chunk of class lines, "first<TAB>id...<TAB>count" each (at least two fields).
struct EqChunk {
    const uint8_t* beg = nullptr; const uint8_t* end = nullptr;
    std::vector<int32_t> ids;
    std::vector<int32_t> sizes;            // ids per class
    std::vector<int64_t> counts;
    std::vector<uint8_t> skip;
    const char* error = nullptr;           // first malformed line of the chunk
};

void parse_eq_chunk(EqChunk& c, int64_t n_contigs) {
    const uint8_t* p = c.beg;
    const size_t guess = (size_t)(c.end - c.beg) / 24 + 16;   // ~30 bytes per class in salmon's files
    c.sizes.reserve(guess); c.counts.reserve(guess); c.skip.reserve(guess); c.ids.reserve(guess * 4);
    while (p < c.end) {
        const uint8_t* nl = static_cast<const uint8_t*>(memchr(p, '\n', (size_t)(c.end - p)));
        const uint8_t* e = nl ? nl : c.end;
        // first token: only "is it exactly 1" matters (read_graph.py:102)
        const uint8_t* t = static_cast<const uint8_t*>(memchr(p, '\t', (size_t)(e - p)));
        if (!t) { c.error = "malformed equivalence class line"; return; }
        const uint8_t skip = (t - p == 1 && *p == '1') ? 1 : 0;
        // the remaining tokens are integers; the last one is the read count, the others contig ids
        int32_t n_ids = 0;
        bool have = false;
        int64_t pending = 0;
        const uint8_t* s = t + 1;
        for (;;) {
            const uint8_t* t2 = static_cast<const uint8_t*>(memchr(s, '\t', (size_t)(e - s)));
            const uint8_t* te = t2 ? t2 : e;
            if (have) {                                        // the previous token was not the last: a contig id
                if (pending < 0 || pending >= n_contigs) { c.error = "contig id out of range"; return; }
                c.ids.push_back((int32_t)pending);
                ++n_ids;
            }
            if (!parse_int(s, te, &pending)) { c.error = t2 ? "contig id out of range" : "class count is not an integer"; return; }
            have = true;
            if (!t2) break;
            s = t2 + 1;
        }
        c.sizes.push_back(n_ids);
        c.counts.push_back(pending);
        c.skip.push_back(skip);
        p = nl ? nl + 1 : c.end;
    }
}

}  // namespace

extern "C" int kb_eq_open(const char* path, kb_eq** out, int64_t* n_contigs, int64_t* n_classes, int64_t* n_ids,
                          int64_t* name_bytes) {
    KB_CHECK_ARG(path && out, "null pointer");
    *out = nullptr;
    uint8_t* img = nullptr;
    int64_t sz = 0;
    { const int rc = kb_host_read_file(path, &img, &sz); if (rc) return rc; }
    struct Guard { uint8_t* p; ~Guard() { free(p); } } guard{img};
    kb_eq* q = new kb_eq();
    const uint8_t* p = img;
    const uint8_t* end = img + sz;
    auto next_line = [&](const uint8_t*& b, const uint8_t*& e) -> bool {
        if (p >= end) return false;
        b = p;
        const uint8_t* nl = static_cast<const uint8_t*>(memchr(p, '\n', (size_t)(end - p)));
        e = nl ? nl : end;
        p = nl ? nl + 1 : end;
        return true;
    };
    const uint8_t *b, *e;
    int64_t n = 0;
    if (!next_line(b, e) || !parse_int(b, e, &n) || n < 0) { delete q; kb_set_error("%s: first line is not the contig count", path); return KB_EINVAL; }
    next_line(b, e);                                            // number of classes: ignored like the reference does
    q->name_off.reserve((size_t)n + 1);
    q->name_off.push_back(0);
    for (int64_t i = 0; i < n; ++i) {
        if (!next_line(b, e)) { b = e = end; }                  // readline() past EOF gives ""
        q->names.insert(q->names.end(), b, e);
        q->name_off.push_back((int64_t)q->names.size());
    }
    // the class lines: chunks of >= 1 MB cut after a '\n', one thread each
    int64_t want = kb_host_threads();
    const int64_t body = (int64_t)(end - p);
    const int64_t min_chunk = getenv("KB_EQ_MIN_CHUNK") ? atoll(getenv("KB_EQ_MIN_CHUNK")) : (1LL << 20);
    if (want > body / (min_chunk > 0 ? min_chunk : 1) + 1) want = body / (min_chunk > 0 ? min_chunk : 1) + 1;
    std::vector<EqChunk> chunks;
    const uint8_t* prev = p;
    for (int64_t t = 1; t <= want; ++t) {
        const uint8_t* cut = t == want ? end : p + body * t / want;
        while (cut < end && cut > p && cut[-1] != '\n') ++cut;
        if (cut > prev) { EqChunk c; c.beg = prev; c.end = cut; chunks.push_back(std::move(c)); prev = cut; }
    }
    if (chunks.size() <= 1) { for (auto& c : chunks) parse_eq_chunk(c, n); }
    else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < chunks.size(); ++i) th.emplace_back([&, i] { parse_eq_chunk(chunks[i], n); });
        for (auto& t : th) t.join();
    }
    size_t tot_classes = 0, tot_ids = 0;
    for (auto& c : chunks) {
        if (c.error) { delete q; kb_set_error("%s: %s", path, c.error); return KB_EINVAL; }
        tot_classes += c.counts.size(); tot_ids += c.ids.size();
    }
    q->class_off.resize(tot_classes + 1);
    q->ids.resize(tot_ids); q->counts.resize(tot_classes); q->skip.resize(tot_classes);
    {
        std::vector<size_t> c0(chunks.size()), i0(chunks.size());
        size_t cc = 0, ii = 0;
        for (size_t i = 0; i < chunks.size(); ++i) { c0[i] = cc; i0[i] = ii; cc += chunks[i].counts.size(); ii += chunks[i].ids.size(); }
        auto place = [&](size_t i) {
            EqChunk& c = chunks[i];
            if (!c.ids.empty()) memcpy(q->ids.data() + i0[i], c.ids.data(), c.ids.size() * 4);
            if (!c.counts.empty()) {
                memcpy(q->counts.data() + c0[i], c.counts.data(), c.counts.size() * 8);
                memcpy(q->skip.data() + c0[i], c.skip.data(), c.skip.size());
            }
            int64_t at = (int64_t)i0[i];
            for (size_t k = 0; k < c.sizes.size(); ++k) { q->class_off[c0[i] + k] = at; at += c.sizes[k]; }
        };
        if (chunks.size() <= 1) { for (size_t i = 0; i < chunks.size(); ++i) place(i); }
        else {
            std::vector<std::thread> th;
            for (size_t i = 0; i < chunks.size(); ++i) th.emplace_back(place, i);
            for (auto& t : th) t.join();
        }
        q->class_off[tot_classes] = (int64_t)tot_ids;
    }
    if (n_contigs) *n_contigs = n;
    if (n_classes) *n_classes = (int64_t)q->counts.size();
    if (n_ids) *n_ids = (int64_t)q->ids.size();
    if (name_bytes) *name_bytes = (int64_t)q->names.size();
    *out = q;
    return KB_OK;
}

extern "C" int kb_eq_fill(kb_eq* q, uint8_t* h_names, int64_t* h_name_off, int64_t* h_class_off, int32_t* h_ids,
                          int64_t* h_counts, uint8_t* h_skip) {
    KB_CHECK_ARG(q, "null handle");
    if (h_names) memcpy(h_names, q->names.data(), q->names.size());
    if (h_name_off) memcpy(h_name_off, q->name_off.data(), q->name_off.size() * 8);
    if (h_class_off) memcpy(h_class_off, q->class_off.data(), q->class_off.size() * 8);
    if (h_ids) memcpy(h_ids, q->ids.data(), q->ids.size() * 4);
    if (h_counts) memcpy(h_counts, q->counts.data(), q->counts.size() * 8);
    if (h_skip) memcpy(h_skip, q->skip.data(), q->skip.size());
    return KB_OK;
}

extern "C" int kb_eq_close(kb_eq* q) { delete q; return KB_OK; }

// ---------------------------------------------------------------------------------------
// device: edges
// ---------------------------------------------------------------------------------------
static void rg_free(kb_ctx* c) {
    cudaFree(c->d_rg_a); cudaFree(c->d_rg_b); cudaFree(c->d_rg_w); cudaFree(c->d_rg_shared);
    c->d_rg_a = c->d_rg_b = nullptr; c->d_rg_w = nullptr; c->d_rg_shared = nullptr; c->rg_edges = 0;
}

extern "C" int kb_readgraph_build(kb_ctx* ctx, int64_t n_contigs, int64_t n_classes, const int64_t* d_class_off,
                                  const int32_t* d_ids, const int64_t* d_counts, const uint8_t* d_skip,
                                  uint64_t* d_totals, int64_t* n_edges) {
    KB_CHECK_ARG(ctx && d_class_off && d_ids && d_counts && d_skip && d_totals && n_edges, "null pointer");
    KB_CHECK_ARG(n_contigs >= 0 && n_contigs < (1LL << 31) && n_classes >= 0, "sizes");
    KB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    { int rc = kb_pool_get(ctx, &DevBuf::pool); if (rc) return rc; }
    DevBuf::stream = st;
    rg_free(ctx);
    *n_edges = 0;
    KB_CUDA(cudaMemsetAsync(d_totals, 0, (size_t)n_contigs * 8, st));
    if (n_classes == 0) return KB_OK;
    KbTimer timer(ctx, 7);
    const unsigned gb = (unsigned)((n_classes + 255) / 256);
    rg_totals<<<gb, 256, 0, st>>>(d_class_off, d_ids, d_counts, n_classes, reinterpret_cast<unsigned long long*>(d_totals));
    ctx->launches++;
    // pairs per class -> exclusive scan
    DevBuf np, po, tmp;
    KB_CUDA(np.alloc((size_t)(n_classes + 1) * 8)); KB_CUDA(po.alloc((size_t)(n_classes + 1) * 8));
    KB_CUDA(cudaMemsetAsync(np.p, 0, (size_t)(n_classes + 1) * 8, st));
    rg_pair_counts<<<gb, 256, 0, st>>>(d_class_off, d_skip, n_classes, np.as<unsigned long long>());
    ctx->launches++;
    size_t tb = 0;
    KB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, np.as<unsigned long long>(), po.as<unsigned long long>(), (int)(n_classes + 1), st));
    KB_CUDA(tmp.alloc(tb));
    KB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, np.as<unsigned long long>(), po.as<unsigned long long>(), (int)(n_classes + 1), st));
    ctx->launches++;
    unsigned long long h_m = 0;
    KB_CUDA(cudaMemcpyAsync(&h_m, po.as<unsigned long long>() + n_classes, 8, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (h_m == 0) return KB_OK;
    if (h_m >= (1ull << 31)) { kb_set_error("read graph: %llu pair occurrences exceed the 2^31 limit of this build", h_m); return KB_EUNSUPPORTED; }
    const int m = (int)h_m;
    // emit + stable sort by (min,max)
    DevBuf k0, k1, s0, s1, c0;
    KB_CUDA(k0.alloc((size_t)m * 8)); KB_CUDA(k1.alloc((size_t)m * 8));
    KB_CUDA(s0.alloc((size_t)m * 4)); KB_CUDA(s1.alloc((size_t)m * 4)); KB_CUDA(c0.alloc((size_t)m * 4));
    const int64_t warps_needed = n_classes;
    const int64_t blocks = (warps_needed * 32 + 255) / 256;
    rg_emit<<<(unsigned)(blocks < 65535 * 4 ? blocks : 65535 * 4), 256, 0, st>>>(d_class_off, d_ids, d_skip, po.as<unsigned long long>(),
                                                                              n_classes, k0.as<uint64_t>(), s0.as<uint32_t>(), c0.as<uint32_t>());
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    size_t sb = 0;
    KB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, k0.as<uint64_t>(), k1.as<uint64_t>(), s0.as<uint32_t>(), s1.as<uint32_t>(), m, 0, 64, st));
    DevBuf tmp2;
    KB_CUDA(tmp2.alloc(sb));
    KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, sb, k0.as<uint64_t>(), k1.as<uint64_t>(), s0.as<uint32_t>(), s1.as<uint32_t>(), m, 0, 64, st));
    ctx->launches++;
    // sorted: keys k1, sequence numbers s1 (ascending inside a run: the sort is stable); class of element = c0[seq]
    // gather class via seq, then heads + counts
    DevBuf head, cnt;
    KB_CUDA(head.alloc((size_t)m)); KB_CUDA(cnt.alloc((size_t)m * 8));
    rg_gather<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(c0.as<uint32_t>(), s1.as<uint32_t>(), m, s0.as<uint32_t>());   // class of every sorted element
    ctx->launches++;
    rg_heads<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(k1.as<uint64_t>(), s0.as<uint32_t>(), d_counts, m, head.as<uint8_t>(),
                                                         cnt.as<unsigned long long>());
    ctx->launches++;
    // reduce runs: unique keys, shared sums, first sequence number
    DevBuf uk, sh, nu, fs;
    KB_CUDA(uk.alloc((size_t)m * 8)); KB_CUDA(sh.alloc((size_t)m * 8)); KB_CUDA(nu.alloc(8)); KB_CUDA(fs.alloc((size_t)m * 4));
    size_t rb = 0;
    KB_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, rb, k1.as<uint64_t>(), uk.as<uint64_t>(), cnt.as<unsigned long long>(),
                                           sh.as<unsigned long long>(), nu.as<int>(), cub::Sum(), m, st));
    DevBuf t4; KB_CUDA(t4.alloc(rb));
    KB_CUDA(cub::DeviceReduce::ReduceByKey(t4.p, rb, k1.as<uint64_t>(), uk.as<uint64_t>(), cnt.as<unsigned long long>(),
                                           sh.as<unsigned long long>(), nu.as<int>(), cub::Sum(), m, st));
    size_t fb = 0;
    KB_CUDA(cub::DeviceSelect::Flagged(nullptr, fb, s1.as<uint32_t>(), head.as<uint8_t>(), fs.as<uint32_t>(), nu.as<int>() + 1, m, st));
    DevBuf t5; KB_CUDA(t5.alloc(fb));
    KB_CUDA(cub::DeviceSelect::Flagged(t5.p, fb, s1.as<uint32_t>(), head.as<uint8_t>(), fs.as<uint32_t>(), nu.as<int>() + 1, m, st));
    ctx->launches += 2;
    int h_u[2] = {0, 0};
    KB_CUDA(cudaMemcpyAsync(h_u, nu.p, 8, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    const int u = h_u[0];
    if (u != h_u[1]) { kb_set_error("internal: run count mismatch (%d vs %d)", h_u[0], h_u[1]); return KB_ECUDA; }
    // order of graph.edges(): by (min_id, first occurrence)
    DevBuf ok0, ok1, oi0, oi1;
    KB_CUDA(ok0.alloc((size_t)u * 8)); KB_CUDA(ok1.alloc((size_t)u * 8)); KB_CUDA(oi0.alloc((size_t)u * 4)); KB_CUDA(oi1.alloc((size_t)u * 4));
    rg_order_keys<<<(unsigned)((u + 255) / 256), 256, 0, st>>>(uk.as<uint64_t>(), fs.as<uint32_t>(), u, ok0.as<uint64_t>(), oi0.as<uint32_t>());
    ctx->launches++;
    size_t ob = 0;
    KB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, ob, ok0.as<uint64_t>(), ok1.as<uint64_t>(), oi0.as<uint32_t>(), oi1.as<uint32_t>(), u, 0, 64, st));
    DevBuf t6; KB_CUDA(t6.alloc(ob));
    KB_CUDA(cub::DeviceRadixSort::SortPairs(t6.p, ob, ok0.as<uint64_t>(), ok1.as<uint64_t>(), oi0.as<uint32_t>(), oi1.as<uint32_t>(), u, 0, 64, st));
    ctx->launches++;
    KB_CUDA(cudaMalloc(&ctx->d_rg_a, (size_t)u * 4)); KB_CUDA(cudaMalloc(&ctx->d_rg_b, (size_t)u * 4));
    KB_CUDA(cudaMalloc(&ctx->d_rg_w, (size_t)u * 8)); KB_CUDA(cudaMalloc(&ctx->d_rg_shared, (size_t)u * 8));
    rg_weights<<<(unsigned)((u + 255) / 256), 256, 0, st>>>(uk.as<uint64_t>(), sh.as<unsigned long long>(), oi1.as<uint32_t>(),
                                                           reinterpret_cast<const unsigned long long*>(d_totals), u, ctx->d_rg_a, ctx->d_rg_b,
                                                           ctx->d_rg_w, reinterpret_cast<unsigned long long*>(ctx->d_rg_shared));
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaStreamSynchronize(st));
    ctx->rg_edges = u;
    *n_edges = u;
    return KB_OK;
}

extern "C" int kb_readgraph_fetch(kb_ctx* ctx, int32_t* h_a, int32_t* h_b, double* h_weight, uint64_t* h_shared) {
    KB_CHECK_ARG(ctx, "ctx");
    const size_t u = (size_t)ctx->rg_edges;
    if (u == 0) return KB_OK;
    KB_CUDA(cudaSetDevice(ctx->device));
    if (h_a) KB_CUDA(cudaMemcpy(h_a, ctx->d_rg_a, u * 4, cudaMemcpyDeviceToHost));
    if (h_b) KB_CUDA(cudaMemcpy(h_b, ctx->d_rg_b, u * 4, cudaMemcpyDeviceToHost));
    if (h_weight) KB_CUDA(cudaMemcpy(h_weight, ctx->d_rg_w, u * 8, cudaMemcpyDeviceToHost));
    if (h_shared) KB_CUDA(cudaMemcpy(h_shared, ctx->d_rg_shared, u * 8, cudaMemcpyDeviceToHost));
    return KB_OK;
}
