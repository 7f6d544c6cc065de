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
#include <cstdlib>

int kb_knn_plan(int sm_count, int impl, int64_t nq, int64_t nk, int32_t k, int64_t n_flag, KbKnnPlan* p) {
    if (k < 1 || nq < 1 || nk < 1 || k > nk) { kb_set_error("kNN: need 1 <= k <= nk and nq >= 1"); return KB_EINVAL; }
    if (k > 24) { kb_set_error("kNN: n_neighbors > 24 not built (candidate lists are <= 32 wide)"); return KB_EUNSUPPORTED; }
    p->impl = impl;
    // candidates kept per (row, split): k plus a margin of >= 6 against fp32 ranking noise, in steps of 8
    p->kp = (k <= 2) ? 8 : (k <= 10 ? 16 : (k <= 18 ? 24 : 32));
    if (impl == KB_KNN_TC) { p->bm = 128; p->bn = 256; }
    else { p->bm = 64; p->bn = 64; }
    p->m_blocks = (nq + p->bm - 1) / p->bm;
    p->n_tiles = (nk + p->bn - 1) / p->bn;
    // Splits: units = (query-block groups) x S are dealt round-robin to the resident CTAs (tensor path:
    // CTA pairs, so sm_count/2 workers and groups of 2 blocks).  Pick S in [1,32] minimising the makespan
    // rounds(S) * tiles_per_unit(S), charging half a tile per unit for list setup and write-back.
    int64_t cl = (impl == KB_KNN_TC && p->m_blocks >= 2) ? 2 : 1;
    if (const char* f = getenv("KB_KNN_CLUSTER")) {             // experiments only: 1, 2 or 4 CTAs share every key tile
        const int v = atoi(f);
        if (impl == KB_KNN_TC && (v == 1 || v == 2 || v == 4) && p->m_blocks >= v) cl = v;
    }
    p->cl = (int)cl;
    const int64_t groups = (p->m_blocks + cl - 1) / cl;
    const int64_t workers = impl == KB_KNN_TC ? (sm_count / cl > 0 ? sm_count / cl : 1) : (int64_t)sm_count * 2;
    int64_t want = 1; double best_cost = 1e300;
    const int64_t max_splits = 512 / p->kp < 32 ? 512 / p->kp : 32;   // K5 merges at most 512 candidates per row
    for (int64_t s_try = 1; s_try <= max_splits && s_try <= p->n_tiles; ++s_try) {
        const int64_t rounds = (groups * s_try + workers - 1) / workers;
        const int64_t per_unit = (p->n_tiles + s_try - 1) / s_try;
        const double cost = (double)rounds * ((double)per_unit + 0.5);
        if (cost < best_cost * 0.995) { best_cost = cost; want = s_try; }
    }
    if (const char* f = getenv("KB_KNN_SPLITS")) {              // experiments only
        const int64_t v = atoll(f);
        if (v >= 1) want = v < p->n_tiles ? v : p->n_tiles;
        if (want > max_splits) want = max_splits;
    }
    p->splits = (int)want;
    p->nk_pad = p->n_tiles * p->bn;
    int64_t off = 0;
    p->off_colmeta = off; off += kb_round_up(p->nk_pad * (int64_t)sizeof(float2), 256);
    p->off_score = off;   off += kb_round_up(nq * p->splits * p->kp * (int64_t)sizeof(float), 256);
    p->off_idx = off;     off += kb_round_up(nq * p->splits * p->kp * (int64_t)sizeof(int32_t), 256);
    p->off_rowthr = off;  off += kb_round_up(nq * (int64_t)sizeof(int32_t), 256);
    p->off_xidx = off;    off += n_flag > 0 ? kb_round_up(nq * p->kp * (int64_t)sizeof(int32_t), 256) : 0;
    p->off_xd2 = off;     off += n_flag > 0 ? kb_round_up(nq * p->kp * (int64_t)sizeof(double), 256) : 0;
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

    KbRowList<KP, BM> list(ls, li);
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
                if (sc < list.bound) list.insert(tid, sc, (int32_t)(n0 + c));
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
                const int32_t* __restrict__ extra_idx, const double* __restrict__ extra_d2,
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
    // ---- 2. exact-side-path extras (rows the tensor path cannot score exactly): lane e also
    //         holds extra candidate e with its exact d2.  A flagged QUERY keeps only extras.
    const int32_t self = (int32_t)(q_row0 + q);
    int32_t x_idx = -1; double x_d2 = 0.0;
    if (extra_idx) {
        if (lane < KP) { x_idx = extra_idx[q * KP + lane]; x_d2 = extra_d2[q * KP + lane]; }
        if (rowmeta[self].flags & 3) my_idx = -1;
    }
    // self must be a candidate: it replaces the last slot if it is missing
    const unsigned has_self = __ballot_sync(0xffffffffu, my_idx == self || x_idx == self);
    if (!has_self && lane == KP - 1) my_idx = self;
    // ---- 3. exact distances: all lanes cooperate on one candidate at a time.
    //   sum_c (a_c*lj - b_c*lq)^2 = lj^2*n_q + lq^2*n_j - 2*lj*lq*g,  g = sum_c a_c*b_c.
    //   Rows that reach this point are unflagged: counts <= 2048 and n < 2^24, hence
    //   g <= sqrt(n_q*n_j) < 2^24 and every partial sum of the fp32 dot product is an
    //   exactly representable integer -- the fp32 FMA chain IS the exact integer Gram entry.
    //   The three terms are exact integers in fp64 (< 2^53), so d2 has one rounding (the division).
    const __half* qrow = op + (int64_t)self * ld;
    const kb_rowmeta mq = rowmeta[self];
    const double lq = (double)mq.key_len;
    double my_d2 = 0.0;
    for (int e = 0; e < KP; ++e) {
        const int32_t j = __shfl_sync(0xffffffffu, my_idx, e);
        if (j < 0 || j == self) continue;                     // d2(self) = 0 exactly
        const __half* krow = op + (int64_t)j * ld;
        float g0 = 0.f, g1 = 0.f;
        for (int c = 8 * lane; c < dp; c += 256) {
            const uint4 a = *reinterpret_cast<const uint4*>(qrow + c);
            const uint4 b = *reinterpret_cast<const uint4*>(krow + c);
            const __half2* ha = reinterpret_cast<const __half2*>(&a);
            const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float2 fa = __half22float2(ha[x]), fb = __half22float2(hb[x]);
                g0 = fmaf(fa.x, fb.x, g0);
                g1 = fmaf(fa.y, fb.y, g1);
            }
        }
        float g = g0 + g1;
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == e) {
            const kb_rowmeta mj = rowmeta[j];
            const double lj = (double)mj.key_len;
            const double num = lj * lj * mq.sqnorm + lq * lq * mj.sqnorm - 2.0 * (lj * lq) * (double)g;
            my_d2 = num / ((lq * lj) * (lq * lj));
        }
    }
    // ---- 4. order: self first, then (d2, idx); rank by counting over both item sets
    const bool valid_a = (lane < KP) && (my_idx >= 0);
    const bool valid_b = (lane < KP) && (x_idx >= 0);
    const double key_a = (my_idx == self) ? -1.0 : my_d2;
    const double key_b = (x_idx == self) ? -1.0 : x_d2;
    int rank_a = 0, rank_b = 0;
    for (int e = 0; e < KP; ++e) {
        const int32_t oj = __shfl_sync(0xffffffffu, my_idx, e);
        const double od = __shfl_sync(0xffffffffu, key_a, e);
        const int32_t xj = __shfl_sync(0xffffffffu, x_idx, e);
        const double xd = __shfl_sync(0xffffffffu, key_b, e);
        if (oj >= 0) {
            if (e != lane && (od < key_a || (od == key_a && oj < my_idx))) ++rank_a;
            if (od < key_b || (od == key_b && oj < x_idx)) ++rank_b;
        }
        if (xj >= 0) {
            if (xd < key_a || (xd == key_a && xj < my_idx)) ++rank_a;
            if (e != lane && (xd < key_b || (xd == key_b && xj < x_idx))) ++rank_b;
        }
    }
    if (valid_a && rank_a < k) {
        out_idx[q * k + rank_a] = my_idx;
        out_dist[q * k + rank_a] = (float)sqrt(my_d2);
        if (out_d2) out_d2[q * k + rank_a] = my_d2;
    }
    if (valid_b && rank_b < k) {
        out_idx[q * k + rank_b] = x_idx;
        out_dist[q * k + rank_b] = (float)sqrt(x_d2);
        if (out_d2) out_d2[q * k + rank_b] = x_d2;
    }
    // fewer valid candidates than k -> mark the rest
    const int n_valid = __popc(__ballot_sync(0xffffffffu, valid_a)) + __popc(__ballot_sync(0xffffffffu, valid_b));
    if (lane >= n_valid && lane < k) {
        out_idx[q * k + lane] = -1;
        out_dist[q * k + lane] = INF;
        if (out_d2) out_d2[q * k + lane] = (double)INF;
    }
}

// ---------------------------------------------------------------------------
// K4x: exact side path.  Rows whose counts do not fit the tensor path exactly
// (flags bit0: a count > 2048, bit1: sum c^2 >= 2^24 -- contigs beyond ~130 kb, long
// homopolymers) are "flagged": masked out of the Gram kernel and handled here in fp64
// from their true u32 counts.  One CTA per query row:
//   flagged query   -> exact d2 to EVERY key
//   unflagged query -> exact d2 to every FLAGGED key
// and the KP best go to K5 as extra candidates.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int find_slot(const int32_t* __restrict__ rows, int n, int32_t row) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int32_t v = rows[mid];
        if (v == row) return mid;
        if (v < row) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

template <int KP>
__global__ void __launch_bounds__(256)
k4x_exact(const __half* __restrict__ op, int64_t ld, int32_t dp, const kb_rowmeta* __restrict__ rowmeta,
          int64_t nk, int64_t q_row0, int64_t nq,
          const int32_t* __restrict__ flag_rows, const uint32_t* __restrict__ flag_counts, int64_t ld_fc,
          int32_t fc_cols, int32_t n_flag,
          int32_t* __restrict__ extra_idx, double* __restrict__ extra_d2) {
    extern __shared__ __align__(16) uint32_t qrow[];          // dp exact counts of the query row
    __shared__ double wl_d[8][KP];
    __shared__ int32_t wl_i[8][KP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double DINF = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
        const int32_t self = (int32_t)(q_row0 + q);
        const kb_rowmeta mq = rowmeta[self];
        const bool q_flagged = (mq.flags & 3) != 0;
        __syncthreads();
        if (q_flagged) {
            const int slot = find_slot(flag_rows, n_flag, self);
            for (int c = threadIdx.x; c < dp; c += 256)
                qrow[c] = (slot >= 0 && c < fc_cols) ? flag_counts[(int64_t)slot * ld_fc + c] : 0u;
        } else {
            for (int c = threadIdx.x; c < dp; c += 256) qrow[c] = (uint32_t)__half2float(op[(int64_t)self * ld + c]);
        }
        if (lane < KP) { wl_d[warp][lane] = DINF; wl_i[warp][lane] = -1; }
        if (KP > 32 && lane + 32 < KP) { wl_d[warp][lane + 32] = DINF; wl_i[warp][lane + 32] = -1; }
        __syncthreads();
        const double lq = (double)mq.key_len;
        const int64_t n_keys = q_flagged ? nk : (int64_t)n_flag;
        double thr = DINF; int pos = 0;                        // lane 0: worst entry of this warp's list
        for (int64_t t = warp; t < n_keys; t += 8) {
            const int32_t j = q_flagged ? (int32_t)t : flag_rows[t];
            const kb_rowmeta mj = rowmeta[j];
            if (mj.flags & 8) continue;                        // padding row of a multi-rank gather
            const double lj = (double)mj.key_len;
            double acc = 0.0;
            if (mj.flags & 3) {
                const int slot = find_slot(flag_rows, n_flag, j);
                const uint32_t* krow = flag_counts + (int64_t)slot * ld_fc;
                for (int c = lane; c < dp; c += 32) {
                    const double b = (slot >= 0 && c < fc_cols) ? (double)krow[c] : 0.0;
                    const double d = (double)qrow[c] * lj - b * lq;
                    acc = fma(d, d, acc);
                }
            } else {
                const __half* krow = op + (int64_t)j * ld;
                for (int c = 2 * lane; c < dp; c += 64) {
                    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(krow + c));
                    const double d0 = (double)qrow[c] * lj - (double)fb.x * lq;
                    const double d1 = (double)qrow[c + 1] * lj - (double)fb.y * lq;
                    acc = fma(d0, d0, acc);
                    acc = fma(d1, d1, acc);
                }
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                const double d2 = acc / ((lq * lj) * (lq * lj));
                if (d2 < thr || (d2 == thr && wl_i[warp][pos] < 0)) {
                    wl_d[warp][pos] = d2; wl_i[warp][pos] = j;
                    double m = -1.0; int mp = 0;
                    for (int e = 0; e < KP; ++e) {
                        const double x = wl_i[warp][e] < 0 ? DINF : wl_d[warp][e];
                        if (x > m) { m = x; mp = e; }
                    }
                    thr = m; pos = mp;
                }
            }
        }
        __syncthreads();
        // merge the 8 warp lists: KP rounds of arg-min over 8*KP entries by one warp
        if (warp == 0) {
            for (int r = 0; r < KP; ++r) {
                double best = DINF; int32_t bi = 0x7fffffff; int bslot = -1;
                for (int e = lane; e < 8 * KP; e += 32) {
                    const int32_t ii = wl_i[e / KP][e % KP];
                    const double dd = wl_d[e / KP][e % KP];
                    if (ii >= 0 && (dd < best || (dd == best && ii < bi))) { best = dd; bi = ii; bslot = e; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    const int os = __shfl_xor_sync(0xffffffffu, bslot, o);
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; bslot = os; }
                }
                if (lane == 0) {
                    extra_idx[q * KP + r] = (bslot >= 0) ? bi : -1;
                    extra_d2[q * KP + r] = (bslot >= 0) ? best : DINF;
                    if (bslot >= 0) wl_i[bslot / KP][bslot % KP] = -1;
                }
                __syncwarp();
            }
        }
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
               const kb_rowmeta* rowmeta, int64_t q_row0, int64_t nq, int32_t k, uint8_t* ws, bool extras,
               int32_t* d_idx, float* d_dist, double* d_d2) {
    const int64_t grid = (nq + 7) / 8;
    k5_merge_rerank<KP><<<(unsigned)grid, 256, 0, ctx->stream>>>(
        op, ld, dp, rowmeta, q_row0, nq, p.splits, k, reinterpret_cast<const float*>(ws + p.off_score),
        reinterpret_cast<const int32_t*>(ws + p.off_idx),
        extras ? reinterpret_cast<const int32_t*>(ws + p.off_xidx) : nullptr,
        extras ? reinterpret_cast<const double*>(ws + p.off_xd2) : nullptr, d_idx, d_dist, d_d2);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

template <int KP>
int run_exact(kb_ctx* ctx, const KbKnnPlan& p, const __half* op, int64_t ld, int32_t dp, const kb_rowmeta* rowmeta,
              int64_t nk, int64_t q_row0, int64_t nq, const int32_t* flag_rows, const uint32_t* flag_counts,
              int64_t ld_fc, int32_t fc_cols, int32_t n_flag, uint8_t* ws) {
    auto kern = k4x_exact<KP>;
    const size_t smem = (size_t)dp * sizeof(uint32_t);
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = nq < (int64_t)ctx->sm_count * 8 ? nq : (int64_t)ctx->sm_count * 8;
    kern<<<(unsigned)grid, 256, smem, ctx->stream>>>(op, ld, dp, rowmeta, nk, q_row0, nq, flag_rows, flag_counts, ld_fc,
                                                     fc_cols, n_flag, reinterpret_cast<int32_t*>(ws + p.off_xidx),
                                                     reinterpret_cast<double*>(ws + p.off_xd2));
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

}  // namespace

extern "C" int64_t kb_knn_workspace_bytes(int64_t nq, int64_t nk, int32_t k, int impl, int64_t n_flag) {
    KbKnnPlan p;
    int64_t best = 0;
    for (int im = KB_KNN_SIMT; im <= KB_KNN_TC; ++im) {
        if (impl != KB_KNN_AUTO && impl != im) continue;
        int rc = kb_knn_plan(148, im, nq, nk, k, n_flag, &p);
        if (rc) return rc;
        if (p.total > best) best = p.total;
    }
    return best;
}

extern "C" int kb_knn(kb_ctx* ctx, int impl, int32_t k,
                      const void* d_operand, int64_t ld_operand, int32_t d_cols_padded,
                      const kb_rowmeta* d_rowmeta,
                      int64_t nk, int64_t q_row0, int64_t nq,
                      const int32_t* d_flag_rows, const uint32_t* d_flag_counts, int64_t ld_flag_counts,
                      int32_t flag_cols, int64_t n_flag,
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
    KB_CHECK_ARG(n_flag >= 0 && n_flag < (1LL << 31) && (n_flag == 0 || (d_flag_rows && d_flag_counts && ld_flag_counts >= flag_cols)),
                 "flagged-row side inputs");
    int rc = kb_knn_plan(ctx->sm_count, impl, nq, nk, k, n_flag, &p);
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
                case 24: rc = run_simt<24>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, ws); break;
                default: rc = run_simt<32>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, ws); break;
            }
        }
        if (rc) return rc;
    }
    const bool extras = n_flag > 0;
    if (extras) {
        KbTimer t(ctx, 6);
        switch (p.kp) {
            case 8: rc = run_exact<8>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, d_flag_rows, d_flag_counts, ld_flag_counts, flag_cols, (int32_t)n_flag, ws); break;
            case 16: rc = run_exact<16>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, d_flag_rows, d_flag_counts, ld_flag_counts, flag_cols, (int32_t)n_flag, ws); break;
            case 24: rc = run_exact<24>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, d_flag_rows, d_flag_counts, ld_flag_counts, flag_cols, (int32_t)n_flag, ws); break;
            default: rc = run_exact<32>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, nk, q_row0, nq, d_flag_rows, d_flag_counts, ld_flag_counts, flag_cols, (int32_t)n_flag, ws); break;
        }
        if (rc) return rc;
    }
    {
        KbTimer t(ctx, 5);
        switch (p.kp) {
            case 8: rc = run_rerank<8>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, extras, d_idx, d_dist, d_d2); break;
            case 16: rc = run_rerank<16>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, extras, d_idx, d_dist, d_d2); break;
            case 24: rc = run_rerank<24>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, extras, d_idx, d_dist, d_d2); break;
            default: rc = run_rerank<32>(ctx, p, op, ld_operand, d_cols_padded, d_rowmeta, q_row0, nq, k, ws, extras, d_idx, d_dist, d_d2); break;
        }
    }
    return rc;
}
