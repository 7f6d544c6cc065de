// kb_count.cu -- K1: per-contig k-mer counting (sm_100a).
//
// Replaces KmerClustering.__count_kmer_occurence (/root/reference/karma/kmer.py:56-92)
// and the window generator __kmers_of_seq (kmer.py:181-197) for windows made of
// A/C/G/T only.  Windows that contain any other byte ("exotic": kmer.py has no
// alphabet, so 'N', lowercase ... get string-keyed columns of their own) are
// only tallied per contig here; kb_exotic.cu enumerates them.
//
// Layout / algorithm
//   bases   uint8[sumL]  ASCII, contigs back to back          (HBM, read once)
//   counts  u32 [n][ld]  one row per contig                   (HBM, written once)
//   One CTA owns one contig at a time (dynamic queue: atomic counter), with the
//   contig's histogram (cols x u32: 4.3 / 5 / 20 / 64 KB) in shared memory.
//   Each thread takes 16 consecutive window starts: one aligned 128-bit load
//   (+ 64-bit halo) -> SIMD ASCII->2-bit conversion -> 5/6/k-mer codes rolled
//   in registers -> shared-memory atomics.  The row is flushed with coalesced
//   128-bit stores (and the histogram cleared in the same pass).  Column
//   presence bits (kmer.py:146-179 "observed k-mers") ride along in registers
//   and are published once per CTA.
//   Contigs longer than KB_LONG_THRESHOLD are queued and handled by a second
//   kernel that tiles each of them over the whole grid (k-1 halo) and merges
//   the partial histograms with global red.add.
//
// Roofline: HBM.  Algorithmic bytes per contig = L (bases) + 4*cols (row).
#include "kb_common.cuh"

#define KB_LONG_THRESHOLD (1 << 16)   // bases; longer contigs take the split path
#define KB_LONG_CHUNK     (1 << 14)   // window starts per CTA work item (split path)

namespace {

// 4 ASCII bytes -> (8 bits of 2-bit codes, first byte most significant; 4 bad bits)
__device__ __forceinline__ void convert4(uint32_t w, uint32_t& code8, uint32_t& bad4) {
    const uint32_t mA = __vcmpeq4(w, 0x41414141u);
    const uint32_t mC = __vcmpeq4(w, 0x43434343u);
    const uint32_t mG = __vcmpeq4(w, 0x47474747u);
    const uint32_t mT = __vcmpeq4(w, 0x54545454u);
    const uint32_t cb = (mC & 0x01010101u) | (mG & 0x02020202u) | (mT & 0x03030303u);
    const uint32_t bb = ~(mA | mC | mG | mT) & 0x01010101u;
    code8 = (cb * 0x40100401u) >> 24;          // c0<<6 | c1<<4 | c2<<2 | c3
    bad4 = ((bb * 0x08040201u) >> 24) & 0xFu;  // b0<<3 | b1<<2 | b2<<1 | b3
}

template <int KA, int KB, bool PALB>
struct Bins {
    static constexpr int A = 1 << (2 * KA);
    static constexpr int B = (KB == 0) ? 0 : (PALB ? (1 << KB) : (1 << (2 * KB)));  // pal: 4^(KB/2)=2^KB
    static constexpr int TOTAL = A + B;
};

// Accumulate windows starting in [lo, hi) of the contig at absolute byte `beg`
// (length L) into the shared histogram.  Returns this thread's exotic-window tally.
template <int KA, int KB, bool PALB, int THREADS>
__device__ __forceinline__ uint32_t accumulate_range(const uint8_t* __restrict__ bases,
                                                     int64_t beg, int64_t L, int64_t lo, int64_t hi,
                                                     uint32_t* hist) {
    constexpr int KMAX = (KB > KA) ? KB : KA;
    static_assert(KMAX <= 8, "halo of one 64-bit load covers k <= 8 only");
    constexpr int BINS_A = Bins<KA, KB, PALB>::A;
    uint32_t exotic = 0;
    const int64_t end_abs = beg + L;
    const int64_t a0 = (beg + lo) & ~int64_t(15);              // absolute, 16 B aligned
    for (int64_t A = a0 + 16 * (int64_t)threadIdx.x; A < beg + hi; A += 16 * THREADS) {
        // ---- load 24 bytes: positions A .. A+23 (absolute)
        uint4 v = make_uint4(0, 0, 0, 0);
        uint2 h = make_uint2(0, 0);
        if (A < end_abs) v = __ldg(reinterpret_cast<const uint4*>(bases + A));
        if (A + 16 < end_abs) h = __ldg(reinterpret_cast<const uint2*>(bases + A + 16));
        uint32_t c0, c1, c2, c3, c4, c5, b0, b1, b2, b3, b4, b5;
        convert4(v.x, c0, b0); convert4(v.y, c1, b1); convert4(v.z, c2, b2);
        convert4(v.w, c3, b3); convert4(h.x, c4, b4); convert4(h.y, c5, b5);
        // 24 bases, base j at bits [2*(23-j), 2*(23-j)+2); bad bit of base j at bit 23-j
        const uint64_t codes = ((uint64_t)((c0 << 8) | c1) << 32) | (uint64_t)((c2 << 24) | (c3 << 16) | (c4 << 8) | c5);
        const uint32_t bad = (b0 << 20) | (b1 << 16) | (b2 << 12) | (b3 << 8) | (b4 << 4) | b5;
        const int64_t s0 = A - beg;                            // contig position of base 0 (may be < 0)
        // window-start index range [wlo, whi) within this thread's 16
        const int64_t wlo64 = lo - s0;
        const int wlo = wlo64 > 0 ? (int)(wlo64 < 16 ? wlo64 : 16) : 0;
        const int64_t whiA64 = ((hi < L - KA + 1) ? hi : (L - KA + 1)) - s0;
        const int whiA = whiA64 < 0 ? 0 : (whiA64 < 16 ? (int)whiA64 : 16);
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            if (w >= wlo && w < whiA) {
                const uint32_t code = (uint32_t)(codes >> (2 * (24 - KA - w))) & (BINS_A - 1);
                const uint32_t bw = (bad >> (24 - KA - w)) & ((1u << KA) - 1);
                if (bw == 0) atomicAdd(&hist[code], 1u);
                else ++exotic;
            }
        }
        if constexpr (KB > 0) {
            const int64_t whiB64 = ((hi < L - KB + 1) ? hi : (L - KB + 1)) - s0;
            const int whiB = whiB64 < 0 ? 0 : (whiB64 < 16 ? (int)whiB64 : 16);
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                if (w >= wlo && w < whiB) {
                    const uint32_t code = (uint32_t)(codes >> (2 * (24 - KB - w))) & ((1u << (2 * KB)) - 1);
                    const uint32_t bw = (bad >> (24 - KB - w)) & ((1u << KB) - 1);
                    if (bw != 0) { ++exotic; continue; }
                    if constexpr (PALB) {
                        // string palindrome x1x2x3x3x2x1 (kmer.py:46-54): compare digit-reversed halves
                        static_assert(!PALB || KB == 6, "palindromic component implemented for 6-mers");
                        const uint32_t hi6 = code >> 6, lo6 = code & 63u;
                        const uint32_t rev = ((lo6 & 3u) << 4) | (lo6 & 12u) | (lo6 >> 4);
                        if (rev == hi6) atomicAdd(&hist[BINS_A + hi6], 1u);
                    } else {
                        atomicAdd(&hist[BINS_A + code], 1u);
                    }
                }
            }
        }
    }
    return exotic;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
    return t;
}

// scratch[0] = next contig, scratch[1] = number of long contigs, scratch[2..3] = exotic total (u64),
// scratch[4..] = rows of the long contigs
template <int KA, int KB, bool PALB, bool PERMUTE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k1_count(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets, int64_t n,
         uint32_t* __restrict__ counts, int64_t ld, uint32_t* __restrict__ exotic_out,
         uint32_t* __restrict__ presence, int32_t* scratch, const uint16_t* __restrict__ perm) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    constexpr int VEC = (COLS + 3) / 4;                       // uint4 per row (COLS % 4 == 0)
    constexpr int ITERS = (VEC + THREADS - 1) / THREADS;
    static_assert(COLS % 4 == 0, "row flush is 128-bit");
    static_assert(ITERS * 4 <= 64, "presence bits live in one 64-bit register");
    extern __shared__ __align__(16) uint32_t hist[];
    __shared__ uint32_t s_red[THREADS / 32];
    __shared__ int64_t s_row;

    for (int i = threadIdx.x; i < COLS; i += THREADS) hist[i] = 0;
    uint64_t pres = 0;

    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_row = (int64_t)atomicAdd(&scratch[0], 1);
        __syncthreads();
        const int64_t row = s_row;
        if (row >= n) break;
        const int64_t beg = offsets[row];
        const int64_t L = offsets[row + 1] - beg;
        uint4* out = reinterpret_cast<uint4*>(counts + row * ld);
        if (L > KB_LONG_THRESHOLD) {
            // queue for the split kernel; zero the row it will red.add into
            if (threadIdx.x == 0) {
                const int slot = atomicAdd(&scratch[1], 1);
                scratch[4 + slot] = (int32_t)row;
                if (exotic_out) exotic_out[row] = 0;
            }
            for (int i = threadIdx.x; i < VEC; i += THREADS) out[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint32_t ex = accumulate_range<KA, KB, PALB, THREADS>(bases, beg, L, 0, L, hist);
        const uint32_t ex_total = block_sum(ex, s_red);        // contains the barrier after accumulation
        if (threadIdx.x == 0) {
            if (exotic_out) exotic_out[row] = ex_total;
            if (ex_total) atomicAdd(reinterpret_cast<unsigned long long*>(scratch + 2), (unsigned long long)ex_total);
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int i = threadIdx.x + it * THREADS;
            if (i < VEC) {
                uint4 r;
                if constexpr (PERMUTE) {
                    const ushort4 p = *reinterpret_cast<const ushort4*>(perm + 4 * i);
                    r = make_uint4(hist[p.x], hist[p.y], hist[p.z], hist[p.w]);
                    hist[p.x] = 0; hist[p.y] = 0; hist[p.z] = 0; hist[p.w] = 0;
                } else {
                    r = *reinterpret_cast<uint4*>(hist + 4 * i);
                    *reinterpret_cast<uint4*>(hist + 4 * i) = make_uint4(0, 0, 0, 0);
                }
                out[i] = r;
                pres |= (uint64_t)((r.x != 0) | ((r.y != 0) << 1) | ((r.z != 0) << 2) | ((r.w != 0) << 3)) << (4 * it);
            }
        }
    }
    if (presence) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int i = threadIdx.x + it * THREADS;
            if (i < VEC) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((pres >> (4 * it + j)) & 1) presence[4 * i + j] = 1u;
            }
        }
    }
}

// Split path: every CTA walks the long-contig list and takes chunks blockIdx.x, +gridDim.x, ...
template <int KA, int KB, bool PALB, bool PERMUTE, int THREADS>
__global__ void __launch_bounds__(THREADS)
k1_count_long(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets,
              uint32_t* __restrict__ counts, int64_t ld, uint32_t* __restrict__ exotic_out,
              uint32_t* __restrict__ presence, const int32_t* __restrict__ scratch,
              const uint16_t* __restrict__ perm) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    extern __shared__ __align__(16) uint32_t hist[];
    __shared__ uint32_t s_red[THREADS / 32];
    const int n_long = scratch[1];
    if (n_long == 0) return;
    for (int i = threadIdx.x; i < COLS; i += THREADS) hist[i] = 0;
    __syncthreads();
    for (int li = 0; li < n_long; ++li) {
        const int64_t row = scratch[4 + li];
        const int64_t beg = offsets[row];
        const int64_t L = offsets[row + 1] - beg;
        const int64_t n_chunks = (L + KB_LONG_CHUNK - 1) / KB_LONG_CHUNK;
        for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const int64_t lo = c * KB_LONG_CHUNK;
            const int64_t hi = (lo + KB_LONG_CHUNK < L) ? lo + KB_LONG_CHUNK : L;
            const uint32_t ex = accumulate_range<KA, KB, PALB, THREADS>(bases, beg, L, lo, hi, hist);
            const uint32_t ex_total = block_sum(ex, s_red);
            if (threadIdx.x == 0 && ex_total) {
                if (exotic_out) atomicAdd(&exotic_out[row], ex_total);
                atomicAdd(reinterpret_cast<unsigned long long*>(const_cast<int32_t*>(scratch) + 2), (unsigned long long)ex_total);
            }
            uint32_t* out = counts + row * ld;
            for (int i = threadIdx.x; i < COLS; i += THREADS) {
                const int src = PERMUTE ? (int)perm[i] : i;
                const uint32_t v = hist[src];
                if (v) {
                    atomicAdd(&out[i], v);
                    hist[src] = 0;
                    if (presence) presence[i] = 1u;
                }
            }
            __syncthreads();
        }
    }
}

template <int KA, int KB, bool PALB, bool PERMUTE, int THREADS>
int launch(kb_ctx* ctx, const uint8_t* d_bases, const int64_t* d_offsets, int64_t n,
           uint32_t* d_counts, int64_t ld, uint32_t* d_exotic, uint32_t* d_presence,
           const uint16_t* d_perm) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    const size_t smem = (size_t)COLS * sizeof(uint32_t);
    auto k1 = k1_count<KA, KB, PALB, PERMUTE, THREADS>;
    auto k1l = k1_count_long<KA, KB, PALB, PERMUTE, THREADS>;
    KB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KB_CUDA(cudaFuncSetAttribute(k1l, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1, THREADS, smem));
    if (per_sm < 1) { kb_set_error("k1_count does not fit on an SM"); return KB_ECUDA; }
    // scratch: counter, n_long, up to n long rows
    const int64_t need = n + 4;
    if (ctx->k1_scratch_cap < need) {
        if (ctx->d_k1_scratch) KB_CUDA(cudaFree(ctx->d_k1_scratch));
        ctx->d_k1_scratch = nullptr; ctx->k1_scratch_cap = 0;
        KB_CUDA(cudaMalloc(&ctx->d_k1_scratch, (size_t)need * sizeof(int32_t)));
        ctx->k1_scratch_cap = need;
    }
    KB_CUDA(cudaMemsetAsync(ctx->d_k1_scratch, 0, 4 * sizeof(int32_t), ctx->stream));
    int64_t grid = (int64_t)ctx->sm_count * per_sm;             // persistent: one resident wave
    if (grid > n) grid = n;
    if (grid < 1) grid = 1;
    {
        KbTimer t(ctx, 0);
        k1<<<(unsigned)grid, THREADS, smem, ctx->stream>>>(d_bases, d_offsets, n, d_counts, ld, d_exotic,
                                                           d_presence, ctx->d_k1_scratch, d_perm);
        ctx->launches++;
    }
    KB_CUDA(cudaGetLastError());
    {
        KbTimer t(ctx, 1);
        k1l<<<(unsigned)(ctx->sm_count * 2), THREADS, smem, ctx->stream>>>(d_bases, d_offsets, d_counts, ld,
                                                                          d_exotic, d_presence,
                                                                          ctx->d_k1_scratch, d_perm);
        ctx->launches++;
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

uint16_t* g_perm_5p6[16] = {nullptr};   // per device

// kmer.py:172-177: sorted() over {5-mers} U {string-palindromic 6-mers}.  For ACGT,
// palindromic 6-mer x1x2x3x3x2x1 sorts directly after its 5-mer prefix x1x2x3x3x2.
int build_perm_5p6(int device, const uint16_t** out) {
    if (device < 0 || device >= 16) { kb_set_error("device index out of range"); return KB_EINVAL; }
    if (!g_perm_5p6[device]) {
        uint16_t h[1088];
        int o = 0;
        for (int c = 0; c < 1024; ++c) {
            h[o++] = (uint16_t)c;
            const int x1 = c >> 8, x2 = (c >> 6) & 3, x3 = (c >> 4) & 3, x4 = (c >> 2) & 3, x5 = c & 3;
            if (x4 == x3 && x5 == x2) h[o++] = (uint16_t)(1024 + (x1 << 4 | x2 << 2 | x3));
        }
        if (o != 1088) { kb_set_error("internal: 5p6 permutation has %d entries", o); return KB_EINVAL; }
        uint16_t* d = nullptr;
        KB_CUDA(cudaMalloc(&d, sizeof(h)));
        KB_CUDA(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
        g_perm_5p6[device] = d;
    }
    *out = g_perm_5p6[device];
    return KB_OK;
}

}  // namespace

int kb_launch_count_kernels(kb_ctx* ctx, const KbMode& m, const uint8_t* d_bases,
                            const int64_t* d_offsets, int64_t n, uint32_t* d_counts,
                            int64_t ld, uint32_t* d_exotic, uint32_t* d_presence) {
#define KB_ARGS ctx, d_bases, d_offsets, n, d_counts, ld, d_exotic, d_presence
    if (m.permute) {
        const uint16_t* perm = nullptr;
        int rc = build_perm_5p6(ctx->device, &perm);
        if (rc) return rc;
        return launch<5, 6, true, true, 128>(KB_ARGS, perm);
    }
    if (m.ka == 5 && m.kb == 6) return launch<5, 6, false, false, 128>(KB_ARGS, nullptr);
    if (m.ka == 4 && m.kb == 5) return launch<4, 5, false, false, 128>(KB_ARGS, nullptr);
    if (m.kb == 0) {
        switch (m.ka) {
            case 1: return launch<1, 0, false, false, 128>(KB_ARGS, nullptr);
            case 2: return launch<2, 0, false, false, 128>(KB_ARGS, nullptr);
            case 3: return launch<3, 0, false, false, 128>(KB_ARGS, nullptr);
            case 4: return launch<4, 0, false, false, 128>(KB_ARGS, nullptr);
            case 5: return launch<5, 0, false, false, 128>(KB_ARGS, nullptr);
            case 6: return launch<6, 0, false, false, 128>(KB_ARGS, nullptr);
            case 7: return launch<7, 0, false, false, 256>(KB_ARGS, nullptr);
        }
    }
#undef KB_ARGS
    kb_set_error("count mode not built");
    return KB_EUNSUPPORTED;
}
