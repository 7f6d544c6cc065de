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
//   Persistent grid, dynamic queue (one atomic counter).  One WARP owns one contig when
//   the histogram is <= 8 KB (5p6, 4+5, k <= 5), one CTA when it is 16-64 KB (5+6, k = 6, 7);
//   the histogram (cols x u32) lives in shared memory.  Each lane takes 16 consecutive
//   window starts: one aligned 128-bit load (+ 64-bit halo) -> bit-parallel ASCII->2-bit
//   conversion and validation -> 5/6/k-mer codes shifted out of a 48-bit register pair ->
//   branch-free shared-memory reductions (windows that must not count go to a dummy word).
//   The string-palindrome test of the 16 six-windows is three XOR/shift masks.  The row is
//   flushed with coalesced 128-bit streaming stores (histogram cleared in the same pass);
//   column presence bits (kmer.py:146-179 "observed k-mers") ride along in a register and
//   are published once per warp/CTA.
//   Contigs longer than the LongPolicy threshold are queued on the device and handled by a
//   second kernel that tiles each of them over the whole grid (k-1 halo) and merges the
//   partial histograms with global red.add.
//
// Roofline: HBM.  Algorithmic bytes per contig = L (bases) + 4*cols (row).  Measured: 0.66-0.69
// of the HBM peak at 5120 columns; at 1088 columns the integer ALU pipe binds first (0.24).
#include "kb_common.cuh"

// Contigs longer than `long_threshold` bases take the split path, `long_chunk` window starts per CTA
// work item.  Large inputs: 64 kb / 16 kb.  Small inputs (fewer contigs than resident warps, e.g. one
// rank's shard of a strong-scaled run): 4 kb / 4 kb, so that one 15 kb contig does not become the tail.
struct LongPolicy { int64_t threshold, chunk; };

namespace {

constexpr unsigned FULL = 0xffffffffu;

// 4 ASCII bytes -> 8 bits of 2-bit codes (first byte most significant).  `xinv` gets a
// non-zero byte wherever the input byte is not one of A/C/G/T.
//   code = ((b>>1) ^ (b>>2)) & 3   maps A,C,G,T -> 0,1,2,3 (and garbage for other bytes),
//   PRMT looks the code up in "ACGT" and the XOR with the input exposes every other byte.
__device__ __forceinline__ uint32_t codes4(uint32_t w, uint32_t& xinv) {
    const uint32_t t = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    const uint32_t sel = (t & 0x3u) | ((t >> 4) & 0x30u) | ((t >> 8) & 0x300u) | ((t >> 12) & 0x3000u);
    xinv = __byte_perm(0x54474341u, 0u, sel) ^ w;
    return (t * 0x40100401u) >> 24;                            // c0<<6 | c1<<4 | c2<<2 | c3
}
// per-byte "non-zero" -> 4 bits, first byte most significant
__device__ __forceinline__ uint32_t badbits4(uint32_t x) {
    const uint32_t y = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
    return (((y >> 7) * 0x08040201u) >> 24) & 0xFu;
}

// Branch-free conditional increment: windows that must not count are steered to a
// per-thread dummy word behind the histogram (ptxas turns a predicated ATOMS into a
// branch; a select + unconditional reduction is shorter and never diverges).
__device__ __forceinline__ void red_shared_inc_if(uint32_t saddr, uint32_t dummy, uint32_t pred) {
    const uint32_t a = pred ? saddr : dummy;
#if defined(KB_K1_EXPERIMENT) && KB_K1_EXPERIMENT == 1      // cost model only (WRONG counts): plain store instead of the reduction
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(pred) : "memory");
#elif defined(KB_K1_EXPERIMENT) && KB_K1_EXPERIMENT == 2    // cost model only (WRONG counts): no shared-memory traffic at all
    asm volatile("" ::"r"(a));
#else
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
#endif
}

template <int KA, int KB, bool PALB>
struct Bins {
    static constexpr int A = 1 << (2 * KA);
    static constexpr int B = (KB == 0) ? 0 : (PALB ? (1 << KB) : (1 << (2 * KB)));  // pal: 4^(KB/2)=2^KB
    static constexpr int TOTAL = A + B;
};

// kmer.py:172-177 column order of "5p6" (sorted(): a palindromic 6-mer follows its 5-mer prefix):
//   column(5-mer code c) = c + #{palindromic 6-mers sorting before it}
//   column(palindromic 6-mer of rank r = 16*x1+4*x2+x3) = 272*x1 + 69*x2 + 21*x3 + 1
// (closed forms checked against sorted() in tests/test_oracle_golden.py::test_full_column_set_ka6).
__device__ __forceinline__ uint32_t sorted_col_5mer(uint32_t c) {
    const uint32_t d1 = c >> 8, d2 = (c >> 6) & 3u;
    const int cc = (int)(c & 63u) - (int)d2;
    const uint32_t t = cc <= 0 ? 0u : (uint32_t)min(4, (cc + 19) / 20);
    return c + 16u * d1 + 4u * d2 + t;
}
__device__ __forceinline__ uint32_t sorted_col_pal6(uint32_t r) {
    return 272u * (r >> 4) + 69u * ((r >> 2) & 3u) + 21u * (r & 3u) + 1u;
}

// Warp-cooperative accumulation of the windows starting in [lo, hi) of the contig at
// absolute byte `beg` (length L) into `hist`.  The warp walks 512-byte spans starting
// at the 16-byte aligned absolute address A_first, stepping A_stride; lane l owns the
// 16 window starts of bytes [A0+16l, A0+16l+16) and loads 8 halo bytes (an L1 hit: they
// are the next lane's first bases).  SORTED: the histogram is kept in kmer.py's sorted
// column order through a byte-offset table (`lut`, 1024 x u16 in shared memory), so the
// row flush is a straight copy.  Returns the lane's tally of windows with a non-ACGT byte.
template <int KA, int KB, bool PALB, bool SORTED>
__device__ __forceinline__ uint32_t accumulate_warp(const uint8_t* __restrict__ bases, int64_t beg, int64_t L,
                                                    int64_t lo, int64_t hi, uint32_t* hist, uint32_t* dummy,
                                                    const uint16_t* lut,
                                                    int64_t A_first, int64_t A_stride, int lane) {
    constexpr int KMAX = (KB > KA) ? KB : KA;
    static_assert(KMAX <= 8, "8 halo bases cover k <= 8 only");
    static_assert(!SORTED || (KA == 5 && KB == 6 && PALB), "sorted layout is the 5p6 mode");
    constexpr uint32_t BINS_A = Bins<KA, KB, PALB>::A;
    uint32_t exotic = 0;
    const int64_t end_abs = beg + L;
    const uint32_t h32 = (uint32_t)__cvta_generic_to_shared(hist);
    const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dummy);
    const int64_t hiA64 = (hi < L - KA + 1) ? hi : (L - KA + 1);
    const int64_t hiB64 = (KB > 0) ? ((hi < L - KB + 1) ? hi : (L - KB + 1)) : 0;
    for (int64_t A0 = A_first; A0 < beg + hi; A0 += A_stride) {          // warp-uniform trip count
        const int64_t A = A0 + 16 * lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (A < end_abs) v = __ldg(reinterpret_cast<const uint4*>(bases + A));
        uint2 h = make_uint2(0, 0);
        if (A + 16 < end_abs) h = __ldg(reinterpret_cast<const uint2*>(bases + A + 16));   // halo: L1 hit (next lane's vector)
        uint32_t x0, x1, x2, x3, y0, y1;
        const uint32_t own = (codes4(v.x, x0) << 24) | (codes4(v.y, x1) << 16) | (codes4(v.z, x2) << 8) | codes4(v.w, x3);
        const uint32_t halo = (codes4(h.x, y0) << 8) | codes4(h.y, y1);
        // 24 bases: base j at bits [2*(23-j), 2*(23-j)+2)
        const uint64_t codes = ((uint64_t)own << 16) | (uint64_t)halo;
        // non-ACGT bytes are rare: the per-base bad mask is only built when some lane saw one
        uint32_t bad = 0;                                                // base j at bit 23-j
        // (the halo only counts when it was loaded: a contig's last vector has none)
        const bool inv_own = ((x0 | x1 | x2 | x3) != 0) && A < end_abs;
        const bool inv_halo = ((y0 | y1) != 0) && A + 16 < end_abs;
        if (__any_sync(FULL, inv_own || inv_halo)) {
            bad = (badbits4(x0) << 20) | (badbits4(x1) << 16) | (badbits4(x2) << 12) | (badbits4(x3) << 8) |
                  (badbits4(y0) << 4) | badbits4(y1);
        }
        // window range of this lane in 32-bit arithmetic: span-uniform 64-bit differences, clamped, then per lane
        const int64_t S0 = A0 - beg;                                     // contig position of lane 0's base 0 (may be < 0)
        const int64_t c_lo = lo - S0, c_hiA = hiA64 - S0;
        const int dlo = (int)(c_lo < -1 ? -1 : (c_lo > 1024 ? 1024 : c_lo)) - 16 * lane;
        const int dhiA = (int)(c_hiA < -1 ? -1 : (c_hiA > 1024 ? 1024 : c_hiA)) - 16 * lane;
        const int wlo = min(max(dlo, 0), 16);
        // ---- component A: window w <-> bit 23-w
        const int whiA = min(max(dhiA, 0), 16);
        const uint32_t rA = whiA > wlo ? (1u << (24 - wlo)) - (1u << (24 - whiA)) : 0u;
        uint32_t BA = bad;
#pragma unroll
        for (int s = 1; s < KA; ++s) BA |= bad << s;
        const uint32_t okA = rA & ~BA;
        exotic += __popc(rA & BA);
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            const uint32_t code = (uint32_t)(codes >> (2 * (24 - KA - w))) & (BINS_A - 1);
            const uint32_t off = SORTED ? (uint32_t)lut[code] : 4u * code;
            red_shared_inc_if(h32 + off, d32, okA & (1u << (23 - w)));
        }
        if constexpr (KB > 0) {
            const int64_t c_hiB = hiB64 - S0;
            const int dhiB = (int)(c_hiB < -1 ? -1 : (c_hiB > 1024 ? 1024 : c_hiB)) - 16 * lane;
            const int whiB = min(max(dhiB, 0), 16);
            const uint32_t rB = whiB > wlo ? (1u << (24 - wlo)) - (1u << (24 - whiB)) : 0u;
            uint32_t BB = bad;
#pragma unroll
            for (int s = 1; s < KB; ++s) BB |= bad << s;
            const uint32_t okB = rB & ~BB;
            exotic += __popc(rB & BB);
            if constexpr (PALB) {
                // string palindrome x1x2x3x3x2x1 (kmer.py:46-54), all 16 windows at once:
                // field j of X_d is zero iff base j == base j+d; window w is a palindrome iff
                // base w==w+5, w+1==w+4, w+2==w+3.  Result: bit 2*(23-w) set <=> NOT a palindrome.
                static_assert(!PALB || KB == 6, "palindromic component implemented for 6-mers");
                const uint64_t X5 = codes ^ (codes << 10), X3 = codes ^ (codes << 6), X1 = codes ^ (codes << 2);
                const uint64_t np = (X5 | (X5 >> 1)) | ((X3 | (X3 >> 1)) << 2) | ((X1 | (X1 >> 1)) << 4);
                // spread the valid-window bits (bit 23-w) to the 2-bit layout (bit 2*(23-w))
                uint32_t sp = okB >> 8;                                  // window w at bit 15-w
                sp = (sp | (sp << 8)) & 0x00FF00FFu; sp = (sp | (sp << 4)) & 0x0F0F0F0Fu;
                sp = (sp | (sp << 2)) & 0x33333333u; sp = (sp | (sp << 1)) & 0x55555555u;
                uint64_t pm = ~np & ((uint64_t)sp << 16);               // valid palindromic windows (~1/64 of all)
                while (__any_sync(FULL, pm != 0)) {                     // warp-uniform trip count: no divergence
                    const int b = pm ? 63 - __clzll((long long)pm) : 4;
                    const uint32_t r = (uint32_t)(codes >> (b - 4)) & 63u;
                    const uint32_t col = SORTED ? sorted_col_pal6(r) : BINS_A + r;
                    red_shared_inc_if(h32 + 4u * col, d32, pm != 0);
                    pm &= ~(1ull << b);
                }
            } else {
#pragma unroll
                for (int w = 0; w < 16; ++w)
                    red_shared_inc_if(h32 + 4u * (BINS_A + ((uint32_t)(codes >> (2 * (24 - KB - w))) & ((1u << (2 * KB)) - 1))),
                                      d32, okB & (1u << (23 - w)));
            }
        }
    }
    return exotic;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Flush one histogram (group of NT threads, this thread = tid) to a count row with 128-bit
// stores, clear it, and collect presence bits (bit 4*it+j for vector tid+it*NT, element j).
template <int COLS, int NT>
__device__ __forceinline__ void flush_row(uint32_t* hist, uint32_t* __restrict__ row_out, int tid, uint64_t& pres) {
    constexpr int VEC = COLS / 4;
    constexpr int ITERS = (VEC + NT - 1) / NT;
    static_assert(COLS % 4 == 0, "row flush is 128-bit");
    static_assert(ITERS * 4 <= 64, "presence bits live in one 64-bit register");
    uint4* out = reinterpret_cast<uint4*>(row_out);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int i = tid + it * NT;
        if (i < VEC) {
            const uint4 r = *reinterpret_cast<uint4*>(hist + 4 * i);
            *reinterpret_cast<uint4*>(hist + 4 * i) = make_uint4(0, 0, 0, 0);
            __stcs(out + i, r);                                          // streaming: the row is not re-read here
            pres |= (uint64_t)((r.x != 0) | ((r.y != 0) << 1) | ((r.z != 0) << 2) | ((r.w != 0) << 3)) << (4 * it);
        }
    }
}

template <int COLS, int NT>
__device__ __forceinline__ void publish_presence(uint32_t* __restrict__ presence, int tid, uint64_t pres) {
    constexpr int VEC = COLS / 4;
    constexpr int ITERS = (VEC + NT - 1) / NT;
    if (!presence) return;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int i = tid + it * NT;
        if (i < VEC) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((pres >> (4 * it + j)) & 1) presence[4 * i + j] = 1u;
        }
    }
}

// scratch[0] = next contig, scratch[1] = number of long contigs, scratch[2..3] = exotic total (u64),
// scratch[4..] = rows of the long contigs

// ---- one WARP per contig (histogram <= 8 KB): no block barriers, 8 independent warps per CTA
template <int KA, int KB, bool PALB, bool SORTED, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k1_count_warp(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets, int64_t n,
              uint32_t* __restrict__ counts, int64_t ld, uint32_t* __restrict__ exotic_out,
              uint32_t* __restrict__ presence, int32_t* scratch, LongPolicy lp) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    extern __shared__ __align__(16) uint32_t smem_hist[];
    const int lane = threadIdx.x & 31;
    uint16_t* lut = reinterpret_cast<uint16_t*>(smem_hist);            // SORTED: 1024 x u16 byte offsets (2 KB)
    uint32_t* hist = smem_hist + (SORTED ? 512 : 0) + (threadIdx.x >> 5) * (COLS + 32);   // + one dummy word per lane
    if constexpr (SORTED) {
        for (int c = threadIdx.x; c < 1024; c += WARPS * 32) lut[c] = (uint16_t)(4u * sorted_col_5mer((uint32_t)c));
    }
    for (int i = lane; i < COLS; i += 32) hist[i] = 0;
    __syncthreads();
    uint64_t pres = 0;
    while (true) {
        int64_t row = 0;
        if (lane == 0) row = (int64_t)atomicAdd(&scratch[0], 1);
        row = __shfl_sync(FULL, row, 0);
        if (row >= n) break;
        const int64_t beg = offsets[row];
        const int64_t L = offsets[row + 1] - beg;
        uint32_t* out = counts + row * ld;
        if (L > lp.threshold) {
            // queue for the split kernel; zero the row it will red.add into
            if (lane == 0) {
                const int slot = atomicAdd(&scratch[1], 1);
                scratch[4 + slot] = (int32_t)row;
                if (exotic_out) exotic_out[row] = 0;
            }
            for (int i = lane; i < COLS / 4; i += 32) reinterpret_cast<uint4*>(out)[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint32_t ex = accumulate_warp<KA, KB, PALB, SORTED>(bases, beg, L, 0, L, hist, hist + COLS + lane, lut, beg & ~int64_t(15), 512, lane);
        const uint32_t ex_total = warp_sum(ex);
        if (lane == 0) {
            if (exotic_out) exotic_out[row] = ex_total;
            if (ex_total) {
                atomicAdd(reinterpret_cast<unsigned long long*>(scratch + 2), (unsigned long long)ex_total);
                if (presence) presence[COLS] = 1u;               // "some window holds a non-ACGT byte"
            }
        }
        __syncwarp();
        flush_row<COLS, 32>(hist, out, lane, pres);
        __syncwarp();
    }
    publish_presence<COLS, 32>(presence, lane, pres);
}

// ---- one CTA per contig (histograms of 16-64 KB)
template <int KA, int KB, bool PALB, bool SORTED, int THREADS>
__global__ void __launch_bounds__(THREADS)
k1_count_cta(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets, int64_t n,
             uint32_t* __restrict__ counts, int64_t ld, uint32_t* __restrict__ exotic_out,
             uint32_t* __restrict__ presence, int32_t* scratch, LongPolicy lp) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    extern __shared__ __align__(16) uint32_t smem_hist[];
    __shared__ uint32_t s_red[THREADS / 32];
    __shared__ int64_t s_row;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* lut = reinterpret_cast<uint16_t*>(smem_hist);
    uint32_t* hist = smem_hist + (SORTED ? 512 : 0);
    if constexpr (SORTED) {
        for (int c = threadIdx.x; c < 1024; c += THREADS) lut[c] = (uint16_t)(4u * sorted_col_5mer((uint32_t)c));
    }
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
        uint32_t* out = counts + row * ld;
        if (L > lp.threshold) {
            if (threadIdx.x == 0) {
                const int slot = atomicAdd(&scratch[1], 1);
                scratch[4 + slot] = (int32_t)row;
                if (exotic_out) exotic_out[row] = 0;
            }
            for (int i = threadIdx.x; i < COLS / 4; i += THREADS) reinterpret_cast<uint4*>(out)[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint32_t ex = warp_sum(accumulate_warp<KA, KB, PALB, SORTED>(bases, beg, L, 0, L, hist, hist + COLS + threadIdx.x, lut,
                                                                          (beg & ~int64_t(15)) + 512 * warp, 16 * THREADS, lane));
        if (lane == 0) s_red[warp] = ex;
        __syncthreads();                                         // all atomics done, s_red visible
        if (threadIdx.x == 0) {
            uint32_t ex_total = 0;
            for (int i = 0; i < THREADS / 32; ++i) ex_total += s_red[i];
            if (exotic_out) exotic_out[row] = ex_total;
            if (ex_total) {
                atomicAdd(reinterpret_cast<unsigned long long*>(scratch + 2), (unsigned long long)ex_total);
                if (presence) presence[COLS] = 1u;
            }
        }
        flush_row<COLS, THREADS>(hist, out, threadIdx.x, pres);
    }
    publish_presence<COLS, THREADS>(presence, threadIdx.x, pres);
}

// ---- split path: every CTA walks the long-contig list and takes chunks blockIdx.x, +gridDim.x, ...
template <int KA, int KB, bool PALB, bool SORTED, int THREADS>
__global__ void __launch_bounds__(THREADS)
k1_count_long(const uint8_t* __restrict__ bases, const int64_t* __restrict__ offsets,
              uint32_t* __restrict__ counts, int64_t ld, uint32_t* __restrict__ exotic_out,
              uint32_t* __restrict__ presence, int32_t* scratch, LongPolicy lp) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    extern __shared__ __align__(16) uint32_t smem_hist[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_long = scratch[1];
    if (n_long == 0) return;
    uint16_t* lut = reinterpret_cast<uint16_t*>(smem_hist);
    uint32_t* hist = smem_hist + (SORTED ? 512 : 0);
    if constexpr (SORTED) {
        for (int c = threadIdx.x; c < 1024; c += THREADS) lut[c] = (uint16_t)(4u * sorted_col_5mer((uint32_t)c));
    }
    for (int i = threadIdx.x; i < COLS; i += THREADS) hist[i] = 0;
    __syncthreads();
    for (int li = 0; li < n_long; ++li) {
        const int64_t row = scratch[4 + li];
        const int64_t beg = offsets[row];
        const int64_t L = offsets[row + 1] - beg;
        const int64_t n_chunks = (L + lp.chunk - 1) / lp.chunk;
        for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const int64_t lo = c * lp.chunk;
            const int64_t hi = (lo + lp.chunk < L) ? lo + lp.chunk : L;
            const uint32_t ex = warp_sum(accumulate_warp<KA, KB, PALB, SORTED>(bases, beg, L, lo, hi, hist, hist + COLS + threadIdx.x, lut,
                                                                              ((beg + lo) & ~int64_t(15)) + 512 * warp, 16 * THREADS, lane));
            if (lane == 0 && ex) {
                if (exotic_out) atomicAdd(&exotic_out[row], ex);
                atomicAdd(reinterpret_cast<unsigned long long*>(scratch + 2), (unsigned long long)ex);
                if (presence) presence[COLS] = 1u;
            }
            __syncthreads();
            uint32_t* out = counts + row * ld;
            for (int i = threadIdx.x; i < COLS; i += THREADS) {
                const uint32_t v = hist[i];
                if (v) {
                    atomicAdd(&out[i], v);
                    hist[i] = 0;
                    if (presence) presence[i] = 1u;
                }
            }
            __syncthreads();
        }
    }
}

template <int KA, int KB, bool PALB, bool SORTED>
int launch(kb_ctx* ctx, const uint8_t* d_bases, const int64_t* d_offsets, int64_t n,
           uint32_t* d_counts, int64_t ld, uint32_t* d_exotic, uint32_t* d_presence) {
    constexpr int COLS = Bins<KA, KB, PALB>::TOTAL;
    constexpr bool WARP_PER_CONTIG = COLS <= 2048;
    constexpr int WARPS = 8;
    constexpr int THREADS = (COLS > 8192) ? 256 : 128;            // CTA-per-contig / split kernels
    // scratch: counter, n_long, exotic total, up to n long rows
    const int64_t need = n + 4;
    if (ctx->k1_scratch_cap < need) {
        if (ctx->d_k1_scratch) KB_CUDA(cudaFree(ctx->d_k1_scratch));
        ctx->d_k1_scratch = nullptr; ctx->k1_scratch_cap = 0;
        KB_CUDA(cudaMalloc(&ctx->d_k1_scratch, (size_t)need * sizeof(int32_t)));
        ctx->k1_scratch_cap = need;
    }
    KB_CUDA(cudaMemsetAsync(ctx->d_k1_scratch, 0, 4 * sizeof(int32_t), ctx->stream));
    const size_t smem_one = (size_t)(COLS + THREADS) * sizeof(uint32_t) + (SORTED ? 2048 : 0);   // histogram + dummy words (+ column table)
    LongPolicy lp{1 << 16, 1 << 14};
    if constexpr (WARP_PER_CONTIG) {
        auto k1 = k1_count_warp<KA, KB, PALB, SORTED, WARPS>;
        const size_t smem = (size_t)(COLS + 32) * sizeof(uint32_t) * WARPS + (SORTED ? 2048 : 0);
        static int per_sm_cache[16] = {0};                     // per device: attribute set + occupancy known
        int& per_sm = per_sm_cache[ctx->device & 15];
        if (per_sm == 0) {
            KB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1, WARPS * 32, smem));
        }
        if (per_sm < 1) { kb_set_error("k1_count_warp does not fit on an SM"); return KB_ECUDA; }
        int64_t grid = (int64_t)ctx->sm_count * per_sm;         // persistent: one resident wave
        const int64_t want = (n + WARPS - 1) / WARPS;
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
        if (n < 2 * (int64_t)ctx->sm_count * per_sm * WARPS) lp = LongPolicy{4096, 4096};
        KbTimer t(ctx, 0);
        k1<<<(unsigned)grid, WARPS * 32, smem, ctx->stream>>>(d_bases, d_offsets, n, d_counts, ld, d_exotic,
                                                              d_presence, ctx->d_k1_scratch, lp);
        ctx->launches++;
    } else {
        auto k1 = k1_count_cta<KA, KB, PALB, SORTED, THREADS>;
        static int per_sm_cache[16] = {0};
        int& per_sm = per_sm_cache[ctx->device & 15];
        if (per_sm == 0) {
            KB_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_one));
            KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1, THREADS, smem_one));
        }
        if (per_sm < 1) { kb_set_error("k1_count_cta does not fit on an SM"); return KB_ECUDA; }
        int64_t grid = (int64_t)ctx->sm_count * per_sm;
        if (grid > n) grid = n;
        if (grid < 1) grid = 1;
        if (n < 2 * (int64_t)ctx->sm_count * per_sm) lp = LongPolicy{16384, 8192};
        KbTimer t(ctx, 0);
        k1<<<(unsigned)grid, THREADS, smem_one, ctx->stream>>>(d_bases, d_offsets, n, d_counts, ld, d_exotic,
                                                               d_presence, ctx->d_k1_scratch, lp);
        ctx->launches++;
    }
    KB_CUDA(cudaGetLastError());
    {
        auto k1l = k1_count_long<KA, KB, PALB, SORTED, THREADS>;
        static bool attr_set[16] = {false};
        if (!attr_set[ctx->device & 15]) {
            KB_CUDA(cudaFuncSetAttribute(k1l, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_one));
            attr_set[ctx->device & 15] = true;
        }
        KbTimer t(ctx, 1);
        k1l<<<(unsigned)(ctx->sm_count * 2), THREADS, smem_one, ctx->stream>>>(d_bases, d_offsets, d_counts, ld,
                                                                              d_exotic, d_presence,
                                                                              ctx->d_k1_scratch, lp);
        ctx->launches++;
    }
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

}  // namespace

int kb_launch_count_kernels(kb_ctx* ctx, const KbMode& m, const uint8_t* d_bases,
                            const int64_t* d_offsets, int64_t n, uint32_t* d_counts,
                            int64_t ld, uint32_t* d_exotic, uint32_t* d_presence) {
#define KB_ARGS ctx, d_bases, d_offsets, n, d_counts, ld, d_exotic, d_presence
    if (m.permute) return launch<5, 6, true, true>(KB_ARGS);     // kmer.py's sorted() column order
    if (m.ka == 5 && m.kb == 6) return launch<5, 6, false, false>(KB_ARGS);
    if (m.ka == 4 && m.kb == 5) return launch<4, 5, false, false>(KB_ARGS);
    if (m.kb == 0) {
        switch (m.ka) {
            case 1: return launch<1, 0, false, false>(KB_ARGS);
            case 2: return launch<2, 0, false, false>(KB_ARGS);
            case 3: return launch<3, 0, false, false>(KB_ARGS);
            case 4: return launch<4, 0, false, false>(KB_ARGS);
            case 5: return launch<5, 0, false, false>(KB_ARGS);
            case 6: return launch<6, 0, false, false>(KB_ARGS);
            case 7: return launch<7, 0, false, false>(KB_ARGS);
        }
    }
#undef KB_ARGS
    kb_set_error("count mode not built");
    return KB_EUNSUPPORTED;
}
