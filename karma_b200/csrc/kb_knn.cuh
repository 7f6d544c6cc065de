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
    int64_t nk_pad;      // colmeta length (multiple of bn)
    // workspace offsets (bytes)
    int64_t off_colmeta, off_score, off_idx, off_rowthr, total;
};

int kb_knn_plan(int sm_count, int impl, int64_t nq, int64_t nk, int32_t k, KbKnnPlan* p);

// score of key j for query i, up to a per-row constant and positive factor:
//   l_i * d2_ij - n_i/l_i = l_i * n_j/l_j^2 - 2 g_ij / l_j  =  fma(g, cm.x, l_i * cm.y)
// colmeta[j] = { -2/l_j , n_j/l_j^2 }  (+inf in .y masks a key)
__device__ __forceinline__ float kb_score(float g, float2 cm, float li) {
    return fmaf(g, cm.x, li * cm.y);
}

// Per-row running top-KP list in shared memory, entry e of row r at [e*ROWS + r]
// (bank = r % 32: conflict-free for one thread per row).
template <int KP, int ROWS>
struct KbRowList {
    float* s; int32_t* i;
    __device__ __forceinline__ void init(int r) {
#pragma unroll
        for (int e = 0; e < KP; ++e) { s[e * ROWS + r] = __int_as_float(0x7f800000); i[e * ROWS + r] = -1; }
    }
    // replace the current worst (at pos) and rescan for the new worst (cold path: an
    // element beat the running bound).  thr/pos stay in registers.
    __device__ __forceinline__ void insert(int r, float v, int32_t j, float& thr, int& pos) {
        s[pos * ROWS + r] = v; i[pos * ROWS + r] = j;
        float m = s[r]; int mp = 0;
#pragma unroll 4
        for (int e = 1; e < KP; ++e) {
            const float x = s[e * ROWS + r];
            if (x > m) { m = x; mp = e; }
        }
        thr = m; pos = mp;
    }
};

int kb_knn_tc_launch(kb_ctx* ctx, const KbKnnPlan& p, const void* d_operand, int64_t ld_operand,
                     int32_t d_cols_padded, const kb_rowmeta* d_rowmeta, int64_t nk, int64_t q_row0,
                     int64_t nq, uint8_t* ws);
