// kb_knn.cuh -- shared between the SIMT and tcgen05 candidate kernels and the rerank.
#pragma once
#include "kb_common.cuh"

// Candidate-search plan: the N x N score matrix is cut into units
// (query block of BM rows) x (split s of the key tiles); each unit keeps a
// running top-KP per query row and writes it to cand_*[q][s][0..KP).
struct KbKnnPlan {
    int impl;            // KB_KNN_SIMT / KB_KNN_TC
    int kp;              // candidates kept per (row, split): 8 / 16 / 32
    int bm, bn;          // unit tile shape
    int64_t m_blocks;    // ceil(nq / bm)
    int64_t n_tiles;     // ceil(nk / bn)
    int splits;          // S
    int cl;              // tensor path: CTAs per cluster sharing one key-tile stream (1, 2 or 4)
    int64_t nk_pad;      // colmeta length (multiple of bn)
    // workspace offsets (bytes)
    int64_t off_colmeta, off_score, off_idx, off_rowthr, off_xidx, off_xd2, total;
};

int kb_knn_plan(int sm_count, int impl, int64_t nq, int64_t nk, int32_t k, int64_t n_flag, KbKnnPlan* p);

// score of key j for query i, up to a per-row constant and positive factor:
//   l_i * d2_ij - n_i/l_i = l_i * n_j/l_j^2 - 2 g_ij / l_j  =  fma(g, cm.x, l_i * cm.y)
// colmeta[j] = { -2/l_j , n_j/l_j^2 }  (+inf in .y masks a key)
__device__ __forceinline__ float kb_score(float g, float2 cm, float li) {
    return fmaf(g, cm.x, li * cm.y);
}

// Per-row running top-KP list in shared memory, entry e of row r at [e*ROWS + r]
// (bank = r % 32: conflict-free for one thread per row).  The list is unsorted and
// organised as KP/8 groups of 8 slots; the worst entry of every group (value + slot) is
// cached in registers, so replacing the overall worst entry costs one 8-slot rescan
// whatever KP is.  `bound` = the KP-th best score seen so far (+inf until the list is full).
template <int KP, int ROWS>
struct KbRowList {
    static constexpr int NG = KP / 8;
    static_assert(KP % 8 == 0 && NG >= 1 && NG <= 8, "KP must be a multiple of 8, at most 64");
    float* s; int32_t* i;
    float gmax[NG]; int gpos[NG];
    float bound;
    __device__ __forceinline__ KbRowList(float* s_, int32_t* i_) : s(s_), i(i_), bound(0.f) {}
    __device__ __forceinline__ void init(int r) {
        const float inf = __int_as_float(0x7f800000);
#pragma unroll
        for (int e = 0; e < KP; ++e) { s[e * ROWS + r] = inf; i[e * ROWS + r] = -1; }
#pragma unroll
        for (int g = 0; g < NG; ++g) { gmax[g] = inf; gpos[g] = 0; }
        bound = inf;
    }
    // cold path: v beat the bound -> it replaces the current overall worst entry
    __device__ __forceinline__ void insert(int r, float v, int32_t j) {
        int gs = 0; float gm = gmax[0];
#pragma unroll
        for (int g = 1; g < NG; ++g) if (gmax[g] > gm) { gm = gmax[g]; gs = g; }
        int ps = gpos[0];
#pragma unroll
        for (int g = 1; g < NG; ++g) if (g == gs) ps = gpos[g];
        const int base = (gs * 8) * ROWS + r;
        s[base + ps * ROWS] = v; i[base + ps * ROWS] = j;
        float m = s[base]; int mp = 0;
#pragma unroll
        for (int e = 1; e < 8; ++e) {
            const float x = s[base + e * ROWS];
            if (x > m) { m = x; mp = e; }
        }
        float b = m;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            if (g == gs) { gmax[g] = m; gpos[g] = mp; }
            else b = fmaxf(b, gmax[g]);
        }
        bound = b;
    }
};

int kb_knn_tc_launch(kb_ctx* ctx, const KbKnnPlan& p, const void* d_operand, int64_t ld_operand,
                     int32_t d_cols_padded, const kb_rowmeta* d_rowmeta, int64_t nk, int64_t q_row0,
                     int64_t nq, uint8_t* ws);
