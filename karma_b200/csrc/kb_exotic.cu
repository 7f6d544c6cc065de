// kb_exotic.cu -- K1x: windows that contain a byte other than A/C/G/T.
//
// /root/reference/karma/kmer.py has no alphabet: `c[sequence[i:i+k]] += 1`
// (kmer.py:72-73, :84-85) makes a dictionary key of ANY window, so 'N',
// lowercase or '\r' produce columns of their own, ordered by Python string
// comparison in sorted() (kmer.py:172).  Those windows are rare in real
// assemblies; they are enumerated here, on the GPU, as 63-bit keys
//     key = sum_t (byte_t + 1) << 9*(6-t)       (t < k <= 7, zero padded)
// whose integer order equals the string order (a shorter k-mer that is a prefix
// of a longer one sorts first, exactly as in Python), then reduced to unique
// keys and per-(row,key) counts with CUB sort/scan (a cold side path; the hot
// counting loop never touches it).
#include "kb_common.cuh"
#include <cub/cub.cuh>

namespace {

__device__ __forceinline__ bool is_acgt(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }

// one warp per contig with exotic windows
__global__ void __launch_bounds__(256)
kx_emit(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets, int64_t n,
        const uint32_t* __restrict__ exotic, int ka, int kb, int pal_b,
        uint64_t* __restrict__ keys, int32_t* __restrict__ rows, unsigned long long* counter,
        unsigned long long capacity) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp; row < n; row += nwarps) {
        if (exotic[row] == 0) continue;
        const uint8_t* s = bases + offsets[row];
        const int64_t L = offsets[row + 1] - offsets[row];
        for (int comp = 0; comp < 2; ++comp) {
            const int k = comp == 0 ? ka : kb;
            if (k == 0) continue;
            const bool pal = comp == 1 && pal_b;
            for (int64_t base = 0; base < L - k + 1; base += 32) {
                const int64_t st = base + lane;
                bool emit = false;
                uint64_t key = 0;
                if (st + k <= L) {
                    bool bad = false, ispal = true;
                    for (int t = 0; t < k; ++t) {
                        const uint8_t b = s[st + t];
                        bad |= !is_acgt(b);
                        ispal &= (b == s[st + k - 1 - t]);
                        key |= (uint64_t)(b + 1u) << (9 * (6 - t));
                    }
                    emit = bad && (!pal || ispal);
                }
                const unsigned m = __ballot_sync(0xffffffffu, emit);
                if (m) {
                    unsigned long long p0 = 0;
                    if (lane == 0) p0 = atomicAdd(counter, (unsigned long long)__popc(m));
                    p0 = __shfl_sync(0xffffffffu, p0, 0);
                    if (emit) {
                        const unsigned long long p = p0 + __popc(m & ((1u << lane) - 1));
                        if (p < capacity) { keys[p] = key; rows[p] = (int32_t)row; }
                    }
                }
            }
        }
    }
}

__global__ void kx_heads(const uint64_t* __restrict__ keys, const int32_t* __restrict__ rows, int64_t m,
                         int32_t* __restrict__ ehead, int32_t* __restrict__ khead) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const bool kh = (i == 0) || keys[i] != keys[i - 1];
    const bool eh = kh || rows[i] != rows[i - 1];
    khead[i] = kh; ehead[i] = eh;
}

__global__ void kx_reduce(const uint64_t* __restrict__ keys, const int32_t* __restrict__ rows, int64_t m,
                          const int32_t* __restrict__ eid, const int32_t* __restrict__ kid,
                          uint64_t* __restrict__ ukeys, int32_t* __restrict__ erow, int32_t* __restrict__ ekey,
                          uint32_t* __restrict__ ecnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t e = eid[i] - 1, k = kid[i] - 1;               // inclusive scans of the head flags
    atomicAdd(&ecnt[e], 1u);
    const bool kh = (i == 0) || keys[i] != keys[i - 1];
    const bool eh = kh || rows[i] != rows[i - 1];
    if (kh) ukeys[k] = keys[i];
    if (eh) { erow[e] = rows[i]; ekey[e] = k; }
}

__global__ void kx_scatter(const int32_t* __restrict__ erow, const int32_t* __restrict__ ekey,
                           const uint32_t* __restrict__ ecnt, int64_t m, const int32_t* __restrict__ key_col,
                           uint32_t* __restrict__ counts, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t col = key_col[ekey[i]];
    if (col >= 0) counts[(int64_t)erow[i] * ld + col] = ecnt[i];
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

extern "C" int kb_exotic_collect(kb_ctx* ctx, int mode, const uint8_t* d_bases, const int64_t* d_offsets, int64_t n,
                                 const uint32_t* d_exotic, int64_t* n_keys, int64_t* n_entries) {
    KB_CHECK_ARG(ctx && d_bases && d_offsets && d_exotic && n_keys && n_entries, "null pointer");
    KbMode md;
    int rc = kb_mode_describe(mode, &md);
    if (rc) return rc;
    KB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // free previous result
    cudaFree(ctx->d_ex_keys); cudaFree(ctx->d_ex_keys_hi); cudaFree(ctx->d_ex_row); cudaFree(ctx->d_ex_keyidx); cudaFree(ctx->d_ex_cnt);
    ctx->d_ex_keys = nullptr; ctx->d_ex_keys_hi = nullptr; ctx->d_ex_row = nullptr; ctx->d_ex_keyidx = nullptr; ctx->d_ex_cnt = nullptr;
    ctx->ex_n_keys = ctx->ex_n_entries = 0;
    *n_keys = 0; *n_entries = 0;
    if (n == 0) return KB_OK;

    // 1. capacity = sum of the per-contig exotic tallies (an upper bound: the tally
    //    counts every exotic 6-window, the emit pass keeps palindromic ones only)
    DevBuf d_total, d_tmp;
    KB_CUDA(d_total.alloc(sizeof(unsigned long long) * 2));
    size_t tmp_bytes = 0;
    KB_CUDA(cub::DeviceReduce::Sum(nullptr, tmp_bytes, d_exotic, d_total.as<unsigned long long>(), (int)n, st));
    KB_CUDA(d_tmp.alloc(tmp_bytes));
    KB_CUDA(cub::DeviceReduce::Sum(d_tmp.p, tmp_bytes, d_exotic, d_total.as<unsigned long long>(), (int)n, st));
    ctx->launches++;
    unsigned long long h_total = 0;
    KB_CUDA(cudaMemcpyAsync(&h_total, d_total.p, sizeof(h_total), cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (h_total == 0) return KB_OK;
    if (h_total >= (1ull << 31)) { kb_set_error("too many non-ACGT windows (%llu) for the exotic side path", h_total); return KB_EUNSUPPORTED; }
    const int64_t cap = (int64_t)h_total;

    // 2. emit (key,row)
    DevBuf k0, k1, r0, r1;
    KB_CUDA(k0.alloc(cap * 8)); KB_CUDA(k1.alloc(cap * 8)); KB_CUDA(r0.alloc(cap * 4)); KB_CUDA(r1.alloc(cap * 4));
    unsigned long long* d_counter = d_total.as<unsigned long long>() + 1;
    KB_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    const int64_t grid = (n + 7) / 8 < (int64_t)ctx->sm_count * 8 ? (n + 7) / 8 : (int64_t)ctx->sm_count * 8;
    kx_emit<<<(unsigned)grid, 256, 0, st>>>(d_bases, d_offsets, n, d_exotic, md.ka, md.kb, md.pal_b,
                                            k0.as<uint64_t>(), r0.as<int32_t>(), d_counter, (unsigned long long)cap);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    unsigned long long h_m = 0;
    KB_CUDA(cudaMemcpyAsync(&h_m, d_counter, sizeof(h_m), cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (h_m > h_total) { kb_set_error("internal: exotic emit overflow (%llu > %llu)", h_m, h_total); return KB_ECUDA; }
    if (h_m == 0) return KB_OK;
    const int m = (int)h_m;

    // 3. sort by (key,row): stable LSD passes, row first then key
    size_t s1 = 0, s2 = 0;
    KB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, s1, r0.as<int32_t>(), r1.as<int32_t>(), k0.as<uint64_t>(), k1.as<uint64_t>(), m, 0, 32, st));
    KB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, s2, k1.as<uint64_t>(), k0.as<uint64_t>(), r1.as<int32_t>(), r0.as<int32_t>(), m, 0, 64, st));
    DevBuf tmp2;
    KB_CUDA(tmp2.alloc(s1 > s2 ? s1 : s2));
    KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, s1, r0.as<int32_t>(), r1.as<int32_t>(), k0.as<uint64_t>(), k1.as<uint64_t>(), m, 0, 32, st));
    KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, s2, k1.as<uint64_t>(), k0.as<uint64_t>(), r1.as<int32_t>(), r0.as<int32_t>(), m, 0, 64, st));
    ctx->launches += 2;
    // sorted: keys in k0, rows in r0

    // 4. heads, scans, reduce
    DevBuf eh, kh, es, ks, tmp3;
    KB_CUDA(eh.alloc((size_t)m * 4)); KB_CUDA(kh.alloc((size_t)m * 4)); KB_CUDA(es.alloc((size_t)m * 4)); KB_CUDA(ks.alloc((size_t)m * 4));
    kx_heads<<<(m + 255) / 256, 256, 0, st>>>(k0.as<uint64_t>(), r0.as<int32_t>(), m, eh.as<int32_t>(), kh.as<int32_t>());
    ctx->launches++;
    size_t s3 = 0;
    KB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, s3, eh.as<int32_t>(), es.as<int32_t>(), m, st));
    KB_CUDA(tmp3.alloc(s3));
    KB_CUDA(cub::DeviceScan::InclusiveSum(tmp3.p, s3, eh.as<int32_t>(), es.as<int32_t>(), m, st));
    KB_CUDA(cub::DeviceScan::InclusiveSum(tmp3.p, s3, kh.as<int32_t>(), ks.as<int32_t>(), m, st));
    ctx->launches += 2;
    int32_t h_ne = 0, h_nk = 0;
    KB_CUDA(cudaMemcpyAsync(&h_ne, es.as<int32_t>() + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaMemcpyAsync(&h_nk, ks.as<int32_t>() + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    KB_CUDA(cudaMalloc(&ctx->d_ex_keys, (size_t)h_nk * 8));
    KB_CUDA(cudaMalloc(&ctx->d_ex_row, (size_t)h_ne * 4));
    KB_CUDA(cudaMalloc(&ctx->d_ex_keyidx, (size_t)h_ne * 4));
    KB_CUDA(cudaMalloc(&ctx->d_ex_cnt, (size_t)h_ne * 4));
    KB_CUDA(cudaMemsetAsync(ctx->d_ex_cnt, 0, (size_t)h_ne * 4, st));
    kx_reduce<<<(m + 255) / 256, 256, 0, st>>>(k0.as<uint64_t>(), r0.as<int32_t>(), m, es.as<int32_t>(), ks.as<int32_t>(),
                                               ctx->d_ex_keys, ctx->d_ex_row, ctx->d_ex_keyidx, ctx->d_ex_cnt);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaStreamSynchronize(st));
    ctx->ex_n_keys = h_nk; ctx->ex_n_entries = h_ne;
    *n_keys = h_nk; *n_entries = h_ne;
    return KB_OK;
}

extern "C" int kb_exotic_fetch(kb_ctx* ctx, uint64_t* h_keys, int32_t* h_entry_row, int32_t* h_entry_key,
                               uint32_t* h_entry_count) {
    KB_CHECK_ARG(ctx, "ctx");
    KB_CUDA(cudaSetDevice(ctx->device));
    if (ctx->ex_n_keys && h_keys)
        KB_CUDA(cudaMemcpy(h_keys, ctx->d_ex_keys, (size_t)ctx->ex_n_keys * 8, cudaMemcpyDeviceToHost));
    if (ctx->ex_n_entries) {
        if (h_entry_row) KB_CUDA(cudaMemcpy(h_entry_row, ctx->d_ex_row, (size_t)ctx->ex_n_entries * 4, cudaMemcpyDeviceToHost));
        if (h_entry_key) KB_CUDA(cudaMemcpy(h_entry_key, ctx->d_ex_keyidx, (size_t)ctx->ex_n_entries * 4, cudaMemcpyDeviceToHost));
        if (h_entry_count) KB_CUDA(cudaMemcpy(h_entry_count, ctx->d_ex_cnt, (size_t)ctx->ex_n_entries * 4, cudaMemcpyDeviceToHost));
    }
    return KB_OK;
}

extern "C" int kb_exotic_scatter(kb_ctx* ctx, const int32_t* d_key_col, uint32_t* d_counts, int64_t ld) {
    KB_CHECK_ARG(ctx && d_key_col && d_counts, "null pointer");
    if (ctx->ex_n_entries == 0) return KB_OK;
    KB_CUDA(cudaSetDevice(ctx->device));
    const int64_t m = ctx->ex_n_entries;
    kx_scatter<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ex_row, ctx->d_ex_keyidx, ctx->d_ex_cnt, m,
                                                                    d_key_col, d_counts, ld);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// =====================================================================================================
// Integer k >= 8 (kmer.py:83-85 accepts any int): 4^k dense bins no longer fit in shared memory (and kmer.py's
// columns are the OBSERVED k-mers anyway), so every window of every contig -- whatever bytes it holds -- becomes a
// 128-bit key (8 bits per character, big-endian, k <= 16: integer order == Python string order of equal-length
// strings), and the same sort / unique / reduce-by-(row, key) machinery as above produces the sorted column keys
// and the (row, column, count) entries.  Results are kept in the context like those of kb_exotic_collect.
// =====================================================================================================
namespace {

__global__ void __launch_bounds__(256)
ks_window_counts(const int64_t* __restrict__ offsets, int64_t n, int k, int64_t* __restrict__ wcount) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t w = 0;
    if (i < n) { w = offsets[i + 1] - offsets[i] - k + 1; if (w < 0) w = 0; }
    wcount[i] = w;
}

// one warp per contig; window w of contig `row` goes to slot woff[row] + w (deterministic, coalesced)
__global__ void __launch_bounds__(256)
ks_emit(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets, int64_t n, int k,
        const int64_t* __restrict__ woff, uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, int32_t* __restrict__ rows,
        int32_t* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp; row < n; row += nwarps) {
        const uint8_t* s = bases + offsets[row];
        const int64_t nw = woff[row + 1] - woff[row];
        const int64_t base = woff[row];
        for (int64_t w = lane; w < nw; w += 32) {
            uint64_t h = 0, l = 0;
            for (int t = 0; t < k; ++t) {
                const uint64_t b = s[w + t];
                if (t < 8) h |= b << (8 * (7 - t)); else l |= b << (8 * (15 - t));
            }
            hi[base + w] = h; lo[base + w] = l; rows[base + w] = (int32_t)row; perm[base + w] = (int32_t)(base + w);
        }
    }
}

template <class T>
__global__ void __launch_bounds__(256)
ks_gather(const T* __restrict__ src, const int32_t* __restrict__ perm, int64_t m, T* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) dst[i] = src[perm[i]];
}

__global__ void ks_heads(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo, const int32_t* __restrict__ rows, int64_t m,
                         int32_t* __restrict__ ehead, int32_t* __restrict__ khead) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const bool kh = (i == 0) || hi[i] != hi[i - 1] || lo[i] != lo[i - 1];
    const bool eh = kh || rows[i] != rows[i - 1];
    khead[i] = kh; ehead[i] = eh;
}

__global__ void ks_reduce(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo, const int32_t* __restrict__ rows, int64_t m,
                          const int32_t* __restrict__ eid, const int32_t* __restrict__ kid,
                          const int32_t* __restrict__ ehead, const int32_t* __restrict__ khead,
                          uint64_t* __restrict__ ukeys_hi, uint64_t* __restrict__ ukeys_lo, int32_t* __restrict__ erow,
                          int32_t* __restrict__ ekey, uint32_t* __restrict__ ecnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t e = eid[i] - 1, k = kid[i] - 1;               // inclusive scans of the head flags
    atomicAdd(&ecnt[e], 1u);
    if (khead[i]) { ukeys_hi[k] = hi[i]; ukeys_lo[k] = lo[i]; }
    if (ehead[i]) { erow[e] = rows[i]; ekey[e] = k; }
}

}  // namespace

extern "C" int kb_kmer_sorted_collect(kb_ctx* ctx, int k, const uint8_t* d_bases, const int64_t* d_offsets, int64_t n,
                                      int64_t* n_keys, int64_t* n_entries) {
    KB_CHECK_ARG(ctx && d_bases && d_offsets && n_keys && n_entries, "null pointer");
    if (k < 1 || k > 16) { kb_set_error("sorted k-mer counting serves 1 <= k <= 16 (got %d)", k); return KB_EUNSUPPORTED; }
    KB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    cudaFree(ctx->d_ex_keys); cudaFree(ctx->d_ex_keys_hi); cudaFree(ctx->d_ex_row); cudaFree(ctx->d_ex_keyidx); cudaFree(ctx->d_ex_cnt);
    ctx->d_ex_keys = nullptr; ctx->d_ex_keys_hi = nullptr; ctx->d_ex_row = nullptr; ctx->d_ex_keyidx = nullptr; ctx->d_ex_cnt = nullptr;
    ctx->ex_n_keys = ctx->ex_n_entries = 0;
    *n_keys = 0; *n_entries = 0;
    if (n == 0) return KB_OK;
    // 1. windows per contig, exclusive scan -> output slots
    DevBuf wc, wo, tmp;
    KB_CUDA(wc.alloc((size_t)(n + 1) * 8)); KB_CUDA(wo.alloc((size_t)(n + 1) * 8));
    ks_window_counts<<<(unsigned)((n + 256) / 256), 256, 0, st>>>(d_offsets, n, k, wc.as<int64_t>());
    ctx->launches++;
    size_t tb = 0;
    KB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, wc.as<int64_t>(), wo.as<int64_t>(), (int)(n + 1), st));
    KB_CUDA(tmp.alloc(tb));
    KB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, wc.as<int64_t>(), wo.as<int64_t>(), (int)(n + 1), st));
    ctx->launches++;
    int64_t h_m = 0;
    KB_CUDA(cudaMemcpyAsync(&h_m, wo.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (h_m == 0) return KB_OK;
    if (h_m >= (1LL << 31)) { kb_set_error("too many k-mer windows (%lld) for the sorted counting path", (long long)h_m); return KB_EUNSUPPORTED; }
    const int64_t m = h_m;
    // 2. emit
    DevBuf hi0, lo0, r0, p0, a64, b64, p1;
    KB_CUDA(hi0.alloc(m * 8)); KB_CUDA(lo0.alloc(m * 8)); KB_CUDA(r0.alloc(m * 4)); KB_CUDA(p0.alloc(m * 4));
    KB_CUDA(a64.alloc(m * 8)); KB_CUDA(b64.alloc(m * 8)); KB_CUDA(p1.alloc(m * 4));
    const int64_t grid = (n + 7) / 8 < (int64_t)ctx->sm_count * 8 ? (n + 7) / 8 : (int64_t)ctx->sm_count * 8;
    ks_emit<<<(unsigned)grid, 256, 0, st>>>(d_bases, d_offsets, n, k, wo.as<int64_t>(), hi0.as<uint64_t>(), lo0.as<uint64_t>(),
                                            r0.as<int32_t>(), p0.as<int32_t>());
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    // 3. stable LSD sort of the index permutation by (hi, lo); the emit order is ascending in the row already
    const unsigned gm = (unsigned)((m + 255) / 256);
    size_t sb = 0;
    KB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, lo0.as<uint64_t>(), a64.as<uint64_t>(), p0.as<int32_t>(), p1.as<int32_t>(), (int)m, 0, 64, st));
    DevBuf tmp2;
    KB_CUDA(tmp2.alloc(sb));
    int32_t* perm = p0.as<int32_t>();
    if (k > 8) {
        KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, sb, lo0.as<uint64_t>(), a64.as<uint64_t>(), p0.as<int32_t>(), p1.as<int32_t>(), (int)m, 8 * (16 - k), 64, st));
        ctx->launches++;
        perm = p1.as<int32_t>();
    }
    // hi permuted by the current order, then the second pass
    ks_gather<uint64_t><<<gm, 256, 0, st>>>(hi0.as<uint64_t>(), perm, m, b64.as<uint64_t>());
    int32_t* perm2 = (perm == p0.as<int32_t>()) ? p1.as<int32_t>() : p0.as<int32_t>();
    KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, sb, b64.as<uint64_t>(), a64.as<uint64_t>(), perm, perm2, (int)m, k < 8 ? 8 * (8 - k) : 0, 64, st));
    ctx->launches += 2;
    // final order: a64 = hi sorted; gather lo and rows through perm2
    DevBuf lo1, r1;
    KB_CUDA(lo1.alloc(m * 8)); KB_CUDA(r1.alloc(m * 4));
    ks_gather<uint64_t><<<gm, 256, 0, st>>>(lo0.as<uint64_t>(), perm2, m, lo1.as<uint64_t>());
    ks_gather<int32_t><<<gm, 256, 0, st>>>(r0.as<int32_t>(), perm2, m, r1.as<int32_t>());
    ctx->launches += 2;
    const uint64_t* s_hi = a64.as<uint64_t>(); const uint64_t* s_lo = lo1.as<uint64_t>(); const int32_t* s_row = r1.as<int32_t>();
    // 4. heads, scans, reduce
    DevBuf eh, kh, es, ks, tmp3;
    KB_CUDA(eh.alloc((size_t)m * 4)); KB_CUDA(kh.alloc((size_t)m * 4)); KB_CUDA(es.alloc((size_t)m * 4)); KB_CUDA(ks.alloc((size_t)m * 4));
    ks_heads<<<gm, 256, 0, st>>>(s_hi, s_lo, s_row, m, eh.as<int32_t>(), kh.as<int32_t>());
    ctx->launches++;
    size_t s3 = 0;
    KB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, s3, eh.as<int32_t>(), es.as<int32_t>(), (int)m, st));
    KB_CUDA(tmp3.alloc(s3));
    KB_CUDA(cub::DeviceScan::InclusiveSum(tmp3.p, s3, eh.as<int32_t>(), es.as<int32_t>(), (int)m, st));
    KB_CUDA(cub::DeviceScan::InclusiveSum(tmp3.p, s3, kh.as<int32_t>(), ks.as<int32_t>(), (int)m, st));
    ctx->launches += 2;
    int32_t h_ne = 0, h_nk = 0;
    KB_CUDA(cudaMemcpyAsync(&h_ne, es.as<int32_t>() + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaMemcpyAsync(&h_nk, ks.as<int32_t>() + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    KB_CUDA(cudaMalloc(&ctx->d_ex_keys, (size_t)h_nk * 8));
    KB_CUDA(cudaMalloc(&ctx->d_ex_keys_hi, (size_t)h_nk * 8));
    KB_CUDA(cudaMalloc(&ctx->d_ex_row, (size_t)h_ne * 4));
    KB_CUDA(cudaMalloc(&ctx->d_ex_keyidx, (size_t)h_ne * 4));
    KB_CUDA(cudaMalloc(&ctx->d_ex_cnt, (size_t)h_ne * 4));
    KB_CUDA(cudaMemsetAsync(ctx->d_ex_cnt, 0, (size_t)h_ne * 4, st));
    ks_reduce<<<gm, 256, 0, st>>>(s_hi, s_lo, s_row, m, es.as<int32_t>(), ks.as<int32_t>(), eh.as<int32_t>(), kh.as<int32_t>(),
                                  ctx->d_ex_keys_hi, ctx->d_ex_keys, ctx->d_ex_row, ctx->d_ex_keyidx, ctx->d_ex_cnt);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaStreamSynchronize(st));
    ctx->ex_n_keys = h_nk; ctx->ex_n_entries = h_ne;
    *n_keys = h_nk; *n_entries = h_ne;
    return KB_OK;
}

extern "C" int kb_kmer_sorted_fetch(kb_ctx* ctx, uint64_t* h_keys_hi, uint64_t* h_keys_lo) {
    KB_CHECK_ARG(ctx, "ctx");
    KB_CUDA(cudaSetDevice(ctx->device));
    if (ctx->ex_n_keys && ctx->d_ex_keys_hi) {
        if (h_keys_hi) KB_CUDA(cudaMemcpy(h_keys_hi, ctx->d_ex_keys_hi, (size_t)ctx->ex_n_keys * 8, cudaMemcpyDeviceToHost));
        if (h_keys_lo) KB_CUDA(cudaMemcpy(h_keys_lo, ctx->d_ex_keys, (size_t)ctx->ex_n_keys * 8, cudaMemcpyDeviceToHost));
    }
    return KB_OK;
}
