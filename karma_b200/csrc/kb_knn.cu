// kb_knn.cu -- kNN driver: plan, key metadata, SIMT candidate kernel (K4-simt),
// candidate merge + exact rerank (K5).  The tcgen05 candidate kernel is in
// kb_knn_tc.cu.
//
// Replaces the neighbour search inside umap.UMAP(...).fit_transform at
// /root/reference/karma/kmer.py:285-290 (euclidean metric, the point itself is
// neighbour 0).  The profile rows are counts/len(key) (kmer.py:120,:213), so
// with integer counts c, n_i = sum c_i^2, g_ij = sum c_i c_j, l = len(key):
//     d2_ij = n_i/l_i^2 + n_j/l_j^2 - 2 g_ij / (l_i l_j)
// K4 gets g_ij exactly (integer Gram, fp32 accumulation below 2^24) and ranks by
// the fp32 norm expansion; K5 recomputes the kept candidates exactly:
//     d2_ij = sum_c (c_ic*l_j - c_jc*l_i)^2 / (l_i*l_j)^2       (fp64, integer terms)
#include "kb_knn.cuh"
#include <math.h>

int kb_knn_plan(int sm_count, int impl, int64_t nq, int64_t nk, int32_t k, KbKnnPlan* p) {
    if (k < 1 || nq < 1 || nk < 1 || k > nk) { kb_set_error("kNN: need 1 <= k <= nk and nq >= 1"); return KB_EINVAL; }
    if (k > 24) { kb_set_error("kNN: n_neighbors > 24 not built (candidate lists are <= 32 wide)"); return KB_EUNSUPPORTED; }
    p->impl = impl;
    p->kp = (k <= 4) ? 8 : (k <= 10 ? 16 : 32);
    if (impl == KB_KNN_TC) { p->bm = 128; p->bn = 256; }
    else { p->bm = 64; p->bn = 64; }
    p->m_blocks = (nq + p->bm - 1) / p->bm;
    p->n_tiles = (nk + p->bn - 1) / p->bn;
    int64_t want = (8LL * sm_count + p->m_blocks - 1) / p->m_blocks;   // >= ~8 units per SM
    if (want > 16) want = 16;
    if (want > p->n_tiles) want = p->n_tiles;
    if (want < 1) want = 1;
    p->splits = (int)want;
    p->nk_pad = p->n_tiles * p->bn;
    int64_t off = 0;
    p->off_colmeta = off; off += kb_round_up(p->nk_pad * (int64_t)sizeof(float2), 256);
    p->off_score = off;   off += kb_round_up(nq * p->splits * p->kp * (int64_t)sizeof(float), 256);
    p->off_idx = off;     off += kb_round_up(nq * p->splits * p->kp * (int64_t)sizeof(int32_t), 256);
    p->off_rowthr = off;  off += kb_round_up(nq * (int64_t)sizeof(int32_t), 256);
    p->total = off;
    return KB_OK;
}

namespace {

__global__ void __launch_bounds__(256)
k4_prep_colmeta(const kb_rowmeta* __restrict__ rowmeta, int64_t nk, int64_t nk_pad, float2* __restrict__ colmeta,
                int32_t* __restrict__ row_thr, int64_t nq) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nq) row_thr[j] = 0x7f800000;                      // +inf as an ordered-int key
    if (j >= nk_pad) return;
    float2 cm;
    kb_rowmeta m;
    m.flags = 3;
    if (j < nk) m = rowmeta[j];
    if (!(m.flags & 3)) {
        const double l = (double)m.key_len;
        cm.x = (float)(-2.0 / l);
        cm.y = (float)(m.sqnorm / (l * l));
    } else {
        cm.x = 0.f;
        cm.y = __int_as_float(0x7f800000);                    // +inf: never a candidate
    }
    colmeta[j] = cm;
}

// ---------------------------------------------------------------------------
// K4-simt: 64x64 score tiles on the CUDA cores (checker / fallback-free small path)
// ---------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256)
k4_simt(const __half* __restrict__ op, int64_t ld, int32_t dp,
        const float2* __restrict__ colmeta, const kb_rowmeta* __restrict__ rowmeta,
        int64_t nk, int64_t q_row0, int64_t nq, int splits, int64_t n_tiles,
        float* __restrict__ cand_score, int32_t* __restrict__ cand_idx) {
    constexpr int BM = 64, BN = 64, BK = 32;
    // operand tiles and the score tile share storage (the score tile is written
    // only after the last k-step of a tile has been consumed)
    __shared__ __align__(16) float ab[2 * BK * (BM + 4)];
    float (*As)[BM + 4] = reinterpret_cast<float (*)[BM + 4]>(ab);
    float (*Bs)[BN + 4] = reinterpret_cast<float (*)[BN + 4]>(ab + BK * (BM + 4));
    float (*tile)[BN + 1] = reinterpret_cast<float (*)[BN + 1]>(ab);
    static_assert(BM * (BN + 1) <= 2 * BK * (BM + 4), "score tile must fit in the operand tiles");
    __shared__ float ls[KP * BM];
    __shared__ int32_t li[KP * BM];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int s = blockIdx.y;
    const int64_t t_lo = n_tiles * s / splits, t_hi = n_tiles * (s + 1) / splits;

    KbRowList<KP, BM> list{ls, li};
    float thr = __int_as_float(0x7f800000); int pos = 0;
    float my_len = 1.f;
    if (tid < BM) {
        list.init(tid);
        const int64_t q = m0 + tid;
        if (q < nq) my_len = (float)rowmeta[q_row0 + q].key_len;
    }
    const int lrow = tid >> 2, lseg = (tid & 3) * 8;
    for (int64_t t = t_lo; t < t_hi; ++t) {
        const int64_t n0 = t * BN;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int k0 = 0; k0 < dp; k0 += BK) {
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (m0 + lrow < nq) va = *reinterpret_cast<const uint4*>(op + (q_row0 + m0 + lrow) * ld + k0 + lseg);
            if (n0 + lrow < nk) vb = *reinterpret_cast<const uint4*>(op + (n0 + lrow) * ld + k0 + lseg);
            const __half2* ha = reinterpret_cast<const __half2*>(&va);
            const __half2* hb = reinterpret_cast<const __half2*>(&vb);
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 fa = __half22float2(ha[e]), fb = __half22float2(hb[e]);
                As[lseg + 2 * e][lrow] = fa.x; As[lseg + 2 * e + 1][lrow] = fa.y;
                Bs[lseg + 2 * e][lrow] = fb.x; Bs[lseg + 2 * e + 1][lrow] = fb.y;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) tile[ty * 4 + x][tx * 4 + y] = acc[x][y];
        __syncthreads();
        if (tid < BM) {
            for (int c = 0; c < BN; ++c) {
                const float sc = kb_score(tile[tid][c], colmeta[n0 + c], my_len);
                if (sc < thr) list.insert(tid, sc, (int32_t)(n0 + c), thr, pos);
            }
        }
        __syncthreads();
    }
    if (tid < BM && m0 + tid < nq) {
        const int64_t base = ((m0 + tid) * splits + s) * KP;
        for (int e = 0; e < KP; ++e) { cand_score[base + e] = ls[e * BM + tid]; cand_idx[base + e] = li[e * BM + tid]; }
    }
}

// ---------------------------------------------------------------------------
// K5: merge the per-split candidate lists by score, rerank exactly, order, emit.
// One warp per query row.
// ---------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256)
k5_merge_rerank(const __half* __restrict__ op, int64_t ld, int32_t dp,
                const kb_rowmeta* __restrict__ rowmeta, int64_t q_row0, int64_t nq, int splits, int32_t k,
                const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                int32_t* __restrict__ out_idx, float* __restrict__ out_dist, double* __restrict__ out_d2) {
    constexpr int MAXC = 16;                                  // splits*KP <= 32*MAXC
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int total = splits * KP;
    const float INF = __int_as_float(0x7f800000);
    // ---- 1. every lane takes candidates lane, lane+32, ... ; KP rounds of warp arg-min
    float cs[MAXC]; int32_t ci[MAXC];
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
        const int e = lane + 32 * u;
        if (e < total) { cs[u] = cand_score[q * total + e]; ci[u] = cand_idx[q * total + e]; }
        else { cs[u] = INF; ci[u] = -1; }
        if (ci[u] < 0) cs[u] = INF;
    }
    int32_t my_idx = -1;                                      // lane e ends up holding merged candidate e
    for (int r = 0; r < KP; ++r) {
        float best = INF; int32_t bidx = 0x7fffffff; int bu = -1;
#pragma unroll
        for (int u = 0; u < MAXC; ++u)
            if (ci[u] >= 0 && (cs[u] < best || (cs[u] == best && ci[u] < bidx))) { best = cs[u]; bidx = ci[u]; bu = u; }
        // warp arg-min over (score, idx)
        float wb = best; int32_t wi = bidx; int wl = lane;
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, wb, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, wi, o);
            const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
            if (ob < wb || (ob == wb && oi < wi)) { wb = ob; wi = oi; wl = ol; }
        }
        if (wi == 0x7fffffff) break;                          // nothing left anywhere
        if (lane == wl) {
#pragma unroll
            for (int u = 0; u < MAXC; ++u) if (u == bu) ci[u] = -1;   // consume
        }
        if (lane == r) my_idx = wi;
    }
    // ---- 2. self must be a candidate: it replaces the last slot if it is missing
    const int32_t self = (int32_t)(q_row0 + q);
    const unsigned has_self = __ballot_sync(0xffffffffu, my_idx == self);
    if (!has_self && lane == KP - 1) my_idx = self;
    // ---- 3. exact distances: all lanes cooperate on one candidate at a time
    const __half* qrow = op + (int64_t)self * ld;
    const double lq = (double)rowmeta[self].key_len;
    double my_d2 = 0.0;
    for (int e = 0; e < KP; ++e) {
        const int32_t j = __shfl_sync(0xffffffffu, my_idx, e);
        if (j < 0) continue;
        const __half* krow = op + (int64_t)j * ld;
        const double lj = (double)rowmeta[j].key_len;
        double acc = 0.0;
        for (int c = 8 * lane; c < dp; c += 256) {
            const uint4 a = *reinterpret_cast<const uint4*>(qrow + c);
            const uint4 b = *reinterpret_cast<const uint4*>(krow + c);
            const __half2* ha = reinterpret_cast<const __half2*>(&a);
            const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float2 fa = __half22float2(ha[x]), fb = __half22float2(hb[x]);
                const double t0 = (double)fa.x * lj - (double)fb.x * lq;
                const double t1 = (double)fa.y * lj - (double)fb.y * lq;
                acc = fma(t0, t0, acc);
                acc = fma(t1, t1, acc);
            }
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double den = (lq * lj) * (lq * lj);
        if (lane == e) my_d2 = acc / den;
    }
    // ---- 4. order: self first, then (d2, idx); rank by counting
    const bool valid = (lane < KP) && (my_idx >= 0);
    const double key_d = (my_idx == self) ? -1.0 : my_d2;
    int rank = 0;
    for (int e = 0; e < KP; ++e) {
        const int32_t oj = __shfl_sync(0xffffffffu, my_idx, e);
        const double od = __shfl_sync(0xffffffffu, key_d, e);
        if (oj >= 0 && e != lane && (od < key_d || (od == key_d && oj < my_idx))) ++rank;
    }
    if (valid && rank < k) {
        out_idx[q * k + rank] = my_idx;
        out_dist[q * k + rank] = (float)sqrt(my_d2);
        if (out_d2) out_d2[q * k + rank] = my_d2;
    }
    // fewer valid candidates than k (only when nk < KP and ... ) -> mark the rest
    const int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
    if (lane >= n_valid && lane < k) {
        out_idx[q * k + lane] = -1;
        out_dist[q * k + lane] = INF;
        if (out_d2) out_d2[q * k + lane] = (double)INF;
    }
}

template <int KP>
int run_simt(kb_ctx* ctx, const KbKnnPlan& p, const __half* op, int64_t ld, int32_t dp,
             const kb_rowmeta* rowmeta, int64_t nk, int64_t q_row0, int64_t nq, uint8_t* ws) {
    dim3 grid((unsigned)p.m_blocks, (unsigned)p.splits);
    k4_simt<KP><<<grid, 256, 0, ctx->stream>>>(op, ld, dp, reinterpret_cast<const float2*>(ws + p.off_colmeta),
                                              rowmeta, nk, q_row0, nq, p.splits, p.n_tiles,
                                              reinterpret_cast<float*>(ws + p.off_score),
                                              reinterpret_cast<int32_t*>(ws + p.off_idx));
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

template <int KP>
int run_rerank(kb_ctx* ctx, const KbKnnPlan& p, const __half* op, int64_t ld, int32_t dp,
               const kb_rowmeta* rowmeta, int64_t q_row0, int64_t nq, int32_t k, uint8_t* ws,
               int32_t* d_idx, float* d_dist, double* d_d2) {
    const int64_t grid = (nq + 7) / 8;
    k5_merge_rerank<KP><<<(unsigned)grid, 256, 0, ctx->stream>>>(
        op, ld, dp, rowmeta, q_row0, nq, p.splits, k, reinterpret_cast<const float*>(ws + p.off_score),
        reinterpret_cast<const int32_t*>(ws + p.off_idx), d_idx, d_dist, d_d2);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

}  // namespace

extern "C" int64_t kb_knn_workspace_bytes(int64_t nq, int64_t nk, int32_t k, int impl) {
    KbKnnPlan p;
    int64_t best = 0;
    for (int im = KB_KNN_SIMT; im <= KB_KNN_TC; ++im) {
        if (impl != KB_KNN_AUTO && impl != im) continue;
        int rc = kb_knn_plan(148, im, nq, nk, k, &p);
        if (rc) return rc;
        if (p.total > best) best = p.total;
    }
    return best;
}

extern "C" int kb_knn(kb_ctx* ctx, int impl, int32_t k,
                      const void* d_operand, int64_t ld_operand, int32_t d_cols_padded,
                      const kb_rowmeta* d_rowmeta,
                      int64_t nk, int64_t q_row0, int64_t nq,
                      int32_t* d_idx, float* d_dist, double* d_d2,
                      void* d_workspace, int64_t workspace_bytes) {
    KB_CHECK_ARG(ctx && d_operand && d_rowmeta && d_idx && d_dist && d_workspace, "null pointer");
    KB_CHECK_ARG(d_cols_padded > 0 && (d_cols_padded % 64) == 0 && ld_operand >= d_cols_padded && (ld_operand % 8) == 0,
                 "operand columns must be padded to a multiple of 64");
    KB_CHECK_ARG(((uintptr_t)d_operand % 16) == 0, "operand must be 16-byte aligned");
    KB_CHECK_ARG(q_row0 >= 0 && nq >= 1 && q_row0 + nq <= nk, "query rows must be a sub-range of the keys");
    KB_CHECK_ARG(nk < (1LL << 31), "more than 2^31 keys");
    if (impl == KB_KNN_AUTO) impl = (nk >= 512) ? KB_KNN_TC : KB_KNN_SIMT;
    KB_CHECK_ARG(impl == KB_KNN_SIMT || impl == KB_KNN_TC, "impl");
    KbKnnPlan p;
    int rc = kb_knn_plan(ctx->sm_count, impl, nq, nk, k, &p);
    if (rc) return rc;
    if (p.splits * p.kp > 32 * 16) { kb_set_error("internal: too many candidates per row"); return KB_EUNSUPPORTED; }
    if (workspace_bytes < p.total) { kb_set_error("kNN workspace: need %lld bytes, got %lld", (long long)p.total, (long long)workspace_bytes); return KB_EWORKSPACE; }
    KB_CHECK_ARG(((uintptr_t)d_workspace % 256) == 0, "workspace must be 256-byte aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
    const __half* op = reinterpret_cast<const __half*>(d_operand);

    k4_prep_colmeta<<<(unsigned)((p.nk_pad + 255) / 256), 256, 0, ctx->stream>>>(
        d_rowmeta, nk, p.nk_pad, reinterpret_cast<float2*>(ws + p.off_colmeta),
        reinterpret_cast<int32_t*>(ws + p.off_rowthr), nq);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    {
        KbTimer t(ctx, 4);
        if (impl == KB_KNN_TC) {
            rc = kb_knn_tc_launch(ctx, p, d_operand, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, ws);
        } else {
            switch (p.kp) {
                case 8: rc = run_simt<8>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, ws); break;
                case 16: rc = run_simt<16>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, ws); break;
                default: rc = run_simt<32>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, ws); break;
            }
        }
        if (rc) return rc;
    }
    {
        KbTimer t(ctx, 5);
        switch (p.kp) {
            case 8: rc = run_rerank<8>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, d_idx, d_dist, d_d2); break;
            case 16: rc = run_rerank<16>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, d_idx, d_dist, d_d2); break;
            default: rc = run_rerank<32>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, d_idx, d_dist, d_d2); break;
        }
    }
    return rc;
}
