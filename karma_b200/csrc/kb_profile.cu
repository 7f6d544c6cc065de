// kb_profile.cu -- K2 (column compaction) and K3 (normalise / convert), sm_100a.
//
// K2 replaces the sorted(set) column dictionary of __extract_kmers
//    (/root/reference/karma/kmer.py:146-179): observed columns only, closed ranks.
// K3 replaces fill_array_for_contig + the scatter loop (kmer.py:108-122, :206-233):
//    profile[i,c] = count / len(key_i) in IEEE fp64 (Python int/int true division
//    of small ints == fp64 division of the two exactly-converted operands), and
//    emits the kNN operand (raw counts as fp16), the exact squared norm and the
//    per-row flags in the same pass over the count row.
// Both are HBM-bound streaming kernels: K3 moves 4*D (read) + 8*D' + 2*Dp (write)
// bytes per contig.
#include "kb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k2_compact(const uint32_t* __restrict__ in, int64_t ld_in, int32_t cols_in,
           const int32_t* __restrict__ colmap, int64_t n,
           uint32_t* __restrict__ out, int64_t ld_out) {
    // one CTA per row chunk; threads stride the source columns (coalesced reads,
    // near-coalesced writes: colmap is monotone for a compaction)
    for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
        const uint32_t* src = in + row * ld_in;
        uint32_t* dst = out + row * ld_out;
        for (int c = threadIdx.x; c < cols_in; c += blockDim.x) {
            const int32_t d = colmap[c];
            if (d >= 0) dst[d] = src[c];
        }
    }
}

// one warp per row.  Rows [n, n_alloc) are written as gather padding.  Dynamic shared memory:
// `cols` bytes of per-CTA column presence (only when `presence` is given).
__global__ void __launch_bounds__(256)
k3_normalise(const uint32_t* __restrict__ counts, int64_t ld, int32_t cols,
             const int32_t* __restrict__ key_len, int64_t n, int64_t n_alloc,
             double* __restrict__ profile, int64_t ld_profile,
             __half* __restrict__ operand, int64_t ld_operand,
             kb_rowmeta* __restrict__ rowmeta, uint32_t* __restrict__ presence, uint32_t* __restrict__ flags_or) {
    extern __shared__ uint8_t s_pres[];
    __shared__ int s_cnt;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int cols4 = cols & ~3;
    if (presence) {
        for (int c = threadIdx.x; c < cols; c += blockDim.x) s_pres[c] = 0;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
    }
    for (int64_t row = warp; row < n_alloc; row += nwarps) {
        if (row >= n) {
            // gather padding: zero counts, never a neighbour
            if (operand)
                for (int64_t c = 8 * lane; c < ld_operand; c += 256)
                    *reinterpret_cast<uint4*>(operand + row * ld_operand + c) = make_uint4(0, 0, 0, 0);
            if (lane == 0 && rowmeta) {
                kb_rowmeta m;
                m.sqnorm = 0.0; m.key_len = 1; m.flags = 11;
                m.cm_x = 0.f; m.cm_y = __int_as_float(0x7f800000);
                m.reserved[0] = m.reserved[1] = 0;
                rowmeta[row] = m;
            }
            continue;
        }
        const uint32_t* src = counts + row * ld;
        const double len = (double)key_len[row];
        unsigned long long sq = 0;
        uint32_t mx = 0;
        // two 128-bit loads in flight per lane before anything is stored
        for (int c = 4 * lane; c < cols4; c += 256) {
            const bool two = c + 128 < cols4;
            const uint4 v0 = __ldcs(reinterpret_cast<const uint4*>(src + c));
            const uint4 v1 = two ? __ldcs(reinterpret_cast<const uint4*>(src + c + 128)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (half == 1 && !two) break;
                const uint4 v = half ? v1 : v0;
                const int cc = c + 128 * half;
                sq += (unsigned long long)v.x * v.x + (unsigned long long)v.y * v.y +
                      (unsigned long long)v.z * v.z + (unsigned long long)v.w * v.w;
                mx = max(max(mx, v.x), max(v.y, max(v.z, v.w)));
                if (presence) {
                    if (v.x) s_pres[cc] = 1;
                    if (v.y) s_pres[cc + 1] = 1;
                    if (v.z) s_pres[cc + 2] = 1;
                    if (v.w) s_pres[cc + 3] = 1;
                }
                if (profile) {
                    double2 a, b;
                    a.x = v.x ? (double)v.x / len : 0.0;
                    a.y = v.y ? (double)v.y / len : 0.0;
                    b.x = v.z ? (double)v.z / len : 0.0;
                    b.y = v.w ? (double)v.w / len : 0.0;
                    if (((ld_profile & 1) == 0)) {
                        double2* p = reinterpret_cast<double2*>(profile + row * ld_profile + cc);
                        __stcs(p, a); __stcs(p + 1, b);
                    } else {
                        double* q = profile + row * ld_profile + cc;
                        q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
                    }
                }
                if (operand) {
                    const __half2 h0 = __floats2half2_rn((float)min(v.x, 2048u), (float)min(v.y, 2048u));
                    const __half2 h1 = __floats2half2_rn((float)min(v.z, 2048u), (float)min(v.w, 2048u));
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(operand + row * ld_operand + cc) = pk;
                }
            }
        }
        // tail columns (cols % 4) and operand zero padding
        for (int c = cols4 + lane; c < cols; c += 32) {
            const uint32_t v = src[c];
            sq += (unsigned long long)v * v;
            mx = max(mx, v);
            if (presence && v) s_pres[c] = 1;
            if (profile) profile[row * ld_profile + c] = v ? (double)v / len : 0.0;
            if (operand) operand[row * ld_operand + c] = __float2half_rn((float)min(v, 2048u));
        }
        if (operand)
            for (int64_t c = cols + lane; c < ld_operand; c += 32) operand[row * ld_operand + c] = __float2half_rn(0.f);
        for (int o = 16; o > 0; o >>= 1) {
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            const int flags = (mx > 2048u ? 1 : 0) | (sq >= (1ull << 24) ? 2 : 0) | (mx == 0 ? 4 : 0);
            if (rowmeta) {
                kb_rowmeta m;
                m.sqnorm = (double)sq;
                m.key_len = key_len[row];
                m.flags = flags;
                if (flags & 3) { m.cm_x = 0.f; m.cm_y = __int_as_float(0x7f800000); }      // K4 never proposes the row
                else { m.cm_x = (float)(-2.0 / len); m.cm_y = (float)((double)sq / (len * len)); }
                m.reserved[0] = m.reserved[1] = 0;
                rowmeta[row] = m;
            }
            if (flags_or && flags) atomicOr(flags_or, (uint32_t)flags);
        }
    }
    if (presence) {
        // publish: one CTA that saw EVERY column proves the dictionary is complete with a single
        // store; only CTAs that did not fall back to per-column stores (checked before writing:
        // a thousand CTAs storing to the same words would cost more than the kernel)
        __syncthreads();
        int mine = 0;
        for (int c = threadIdx.x; c < cols; c += blockDim.x) mine += s_pres[c];
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0 && mine) atomicAdd(&s_cnt, mine);
        __syncthreads();
        if (s_cnt == cols) {
            if (threadIdx.x == 0 && !(__ldcg(presence + cols) & 2u)) atomicOr(presence + cols, 2u);
        } else {
            for (int c = threadIdx.x; c < cols; c += blockDim.x)
                if (s_pres[c] && __ldcg(presence + c) == 0u) presence[c] = 1u;
        }
    }
}

// OR of the flags of rows [0, n) into one word (rows with bit3 = gather padding are skipped)
__global__ void __launch_bounds__(256)
k3_flags_or(const kb_rowmeta* __restrict__ rowmeta, int64_t n, uint32_t* __restrict__ out) {
    uint32_t v = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t f = rowmeta[i].flags;
        if (!(f & 8)) v |= (uint32_t)f;
    }
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicOr(out, v);
}

}  // namespace

extern "C" int kb_rowmeta_flags_or(kb_ctx* ctx, const kb_rowmeta* d_rowmeta, int64_t n, uint32_t* d_out) {
    KB_CHECK_ARG(ctx && d_rowmeta && d_out && n >= 0, "null pointer");
    KB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(uint32_t), ctx->stream));
    if (n == 0) return KB_OK;
    const int64_t blocks = (n + 255) / 256;
    k3_flags_or<<<(unsigned)(blocks < 296 ? blocks : 296), 256, 0, ctx->stream>>>(d_rowmeta, n, d_out);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

extern "C" int kb_compact(kb_ctx* ctx, const uint32_t* d_in, int64_t ld_in, int32_t d_cols_in,
                          const int32_t* d_colmap, int64_t n,
                          uint32_t* d_out, int64_t ld_out, int32_t d_cols_out) {
    KB_CHECK_ARG(ctx && d_in && d_colmap && d_out, "null pointer");
    KB_CHECK_ARG(n >= 0 && d_cols_in > 0 && d_cols_out > 0 && ld_in >= d_cols_in && ld_out >= d_cols_out, "shape");
    if (n == 0) return KB_OK;
    KbTimer t(ctx, 2);
    KB_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n * ld_out * sizeof(uint32_t), ctx->stream));
    const int64_t grid = n < (int64_t)ctx->sm_count * 8 ? n : (int64_t)ctx->sm_count * 8;
    k2_compact<<<(unsigned)grid, 256, 0, ctx->stream>>>(d_in, ld_in, d_cols_in, d_colmap, n, d_out, ld_out);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

extern "C" int kb_normalise(kb_ctx* ctx, const uint32_t* d_counts, int64_t ld, int32_t d_cols,
                            const int32_t* d_key_len, int64_t n, int64_t n_alloc,
                            double* d_profile, int64_t ld_profile,
                            void* d_operand, int64_t ld_operand,
                            kb_rowmeta* d_rowmeta, uint32_t* d_presence, uint32_t* d_flags_or) {
    KB_CHECK_ARG(ctx && (n == 0 || (d_counts && d_key_len)), "null pointer");
    KB_CHECK_ARG(n >= 0 && n_alloc >= n && d_cols > 0 && ld >= d_cols, "shape");
    KB_CHECK_ARG((ld % 4) == 0 && ((uintptr_t)d_counts % 16) == 0, "counts must be 16-byte aligned with ld % 4 == 0");
    KB_CHECK_ARG(!d_profile || (ld_profile >= d_cols && ((uintptr_t)d_profile % 16) == 0), "profile ld/alignment");
    KB_CHECK_ARG(!d_operand || (ld_operand >= d_cols && (ld_operand % 64) == 0 && ((uintptr_t)d_operand % 16) == 0),
                 "operand ld must be a multiple of 64 and >= columns");
    if (n_alloc == 0) return KB_OK;
    KbTimer t(ctx, 3);
    const int64_t blocks_needed = (n_alloc + 7) / 8;
    const int64_t grid = blocks_needed < (int64_t)ctx->sm_count * 8 ? blocks_needed : (int64_t)ctx->sm_count * 8;
    const size_t smem = d_presence ? (size_t)kb_round_up(d_cols, 16) : 0;
    k3_normalise<<<(unsigned)grid, 256, smem, ctx->stream>>>(d_counts, ld, d_cols, d_key_len, n, n_alloc, d_profile, ld_profile,
                                                             reinterpret_cast<__half*>(d_operand), ld_operand,
                                                             d_rowmeta, d_presence, d_flags_or);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}
