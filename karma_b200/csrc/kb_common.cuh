// kb_common.cuh -- shared declarations for libkarma_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <nvtx3/nvToolsExt.h>          // header-only (NVTX v3): ranges are no-ops unless a profiler is attached
#include "../../include/karma_b200.h"

#define KB_N_TIMERS 9
#define KB_EV_RING 128

struct kb_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;
    int timing;
    // per-stage ring of CUDA-event pairs recorded on the launching stream; read (and
    // reset) by kb_stage_ms without synchronising inside the timed region
    cudaEvent_t (*ev0)[KB_EV_RING];
    cudaEvent_t (*ev1)[KB_EV_RING];
    int ev_n[KB_N_TIMERS];          // launches recorded since the last reset
    int64_t launches;
    // K1 scratch: work counter + long-contig list
    int32_t* d_k1_scratch;          // [0] contig counter, [1] n_long, [2..] long ids
    int64_t k1_scratch_cap;         // capacity in int32 entries
    // exotic side path (owned)
    uint64_t* d_ex_keys;            // unique keys (sorted k-mer path: the low 8 characters)
    uint64_t* d_ex_keys_hi;         // sorted k-mer path (k >= 8): the first 8 characters of every unique key
    int32_t* d_ex_row; int32_t* d_ex_keyidx; uint32_t* d_ex_cnt;
    int64_t ex_n_keys, ex_n_entries;
    // read-graph edges of the last kb_readgraph_build (owned)
    int32_t* d_rg_a; int32_t* d_rg_b; double* d_rg_w; uint64_t* d_rg_shared; int64_t rg_edges;
    // group-pair connection weights of the last kb_links_build: one grow-only scratch block (owned), results point into it
    void* d_lk_scratch; int64_t lk_scratch_bytes;
    int32_t* d_lk_a; int32_t* d_lk_b; double* d_lk_w; int64_t* d_lk_edges; int64_t* d_lk_over; int64_t lk_pairs;
    // stream-ordered pool for the temporaries of the read-graph build (created lazily, keeps its memory)
    cudaMemPool_t pool;
    // tensor-map encoder (driver entry point, resolved lazily)
    void* encode_tiled;
    // kNN plans and uploaded piece tables (kb_knn.cu)
    void* knn_cache;
};

void kb_set_error(const char* fmt, ...);
int kb_cuda_fail(cudaError_t e, const char* what);

#define KB_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return kb_cuda_fail(_e, #call);          \
    } while (0)

#define KB_CHECK_ARG(cond, msg)                                         \
    do {                                                                \
        if (!(cond)) { kb_set_error("invalid argument: %s", msg); return KB_EINVAL; } \
    } while (0)

// One stage of the path on the host timeline: an NVTX range (per-stage ranges for nsys / ncu --nvtx) and,
// with timing enabled, a CUDA-event pair on the launching stream.
static const char* const kb_stage_names[KB_N_TIMERS] = {
    "kb:K1 count", "kb:K1 count (split contigs)", "kb:K2 compact", "kb:K3 normalise", "kb:K4 distance GEMM + top-k",
    "kb:K5 merge + exact rerank", "kb:K4x/K6 exact distances", "kb:read graph", "kb:group links"};
struct KbTimer {
    kb_ctx* c; int which; int slot;
    KbTimer(kb_ctx* ctx, int w) : c(ctx), which(w), slot(0) {
        nvtxRangePushA(kb_stage_names[which]);
        if (c->timing) { slot = c->ev_n[which] % KB_EV_RING; cudaEventRecord(c->ev0[which][slot], c->stream); }
    }
    ~KbTimer() {
        if (c->timing) { cudaEventRecord(c->ev1[which][slot], c->stream); c->ev_n[which]++; }
        nvtxRangePop();
    }
};

// mode description shared by host and device
struct KbMode {
    int ka;        // first component k (0 = none)
    int kb;        // second component k (0 = none)
    int pal_b;     // second component: string-palindromic windows only, binned by rank
    int bins_a;    // 4^ka
    int bins_b;    // 4^kb, or 4^(kb/2) when pal_b
    int permute;   // columns are a permutation of [A bins | B bins] (5p6 sorted order)
    int cols;      // bins_a + bins_b
};
int kb_mode_describe(int mode, KbMode* out);

// internal launchers (defined in the per-kernel .cu files)
int kb_launch_count_kernels(kb_ctx* ctx, const KbMode& m, const uint8_t* d_bases,
                            const int64_t* d_offsets, int64_t n, uint32_t* d_counts,
                            int64_t ld, uint32_t* d_exotic, uint32_t* d_presence, int track_columns);

void kb_links_free(kb_ctx* c);
void kb_knn_cache_free(kb_ctx* c);
// host-side file input shared by the FASTA and eq_classes parsers (kb_fasta.cu)
int64_t kb_host_threads();
int kb_host_read_file(const char* path, uint8_t** img, int64_t* size);
int kb_pool_get(kb_ctx* c, cudaMemPool_t* out);

static inline int64_t kb_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
