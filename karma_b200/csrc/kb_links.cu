// kb_links.cu -- connection weights between groups of contigs of the read graph
// (follows the read-graph path, separate from the k-mer front end).
//
// Replaces the Python product loops of
//   calc_connections_between_mcl_subclusters   /root/reference/karma/karma.py:103-118
//   ReadGraph.calc_distance_between_subgraphs  /root/reference/karma/read_graph.py:359-373
// Both walk itertools.product(nodes_A, nodes_B), test has_edge(A, B) and add the edge weight to
// a running float64 sum: O(|A|*|B|) dictionary probes per pair of groups, O(S^2) pairs.
//
// GPU formulation (sort / segmented sequential sum; HBM-bound integer + fp64 work):
//   every node may have a ROW role (group ra, position pa inside that group's list) and a COLUMN
//   role (rb, pb); an undirected edge {u,v} is visited by the product of groups (ra < rb) at
//   most twice: as (u,v) when u has a row role and v a column role with ra[u] < rb[v], and as
//   (v,u) likewise (u != v).  Each visit becomes an item keyed by (ra, rb | pa, pb).
//   1. emit items (warp-aggregated append), 2. radix sort by (pa,pb) then stable by (ra,rb),
//   3. run-length encode the group pairs, 4. one thread per pair adds its weights IN PRODUCT
//   ORDER (row position major, column position minor), which is the reference's summation order,
//   so the float64 total is bit-identical; it also counts the edges at which the running sum
//   exceeds the cut-off, which is how often karma.py:116-117 appends the pair.
// For the sub-cluster partition of karma.py both roles are the node's sub-cluster; for two node
// lists the row role is "in nodes_a" (group 0) and the column role "in nodes_b" (group 1).
#include "kb_common.cuh"
#include <cub/cub.cuh>

namespace {

__device__ __forceinline__ void lk_append(bool valid, uint64_t k1, uint64_t k2, uint32_t edge,
                                          unsigned long long* counter, uint64_t* key1, uint64_t* key2, uint32_t* src) {
    const unsigned mask = __ballot_sync(0xffffffffu, valid);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid) {
        const unsigned long long at = base + __popc(mask & ((1u << lane) - 1));
        key1[at] = k1; key2[at] = k2; src[at] = edge;
    }
}

// one thread per edge; every thread of a warp reaches both appends (the ballots are warp-wide)
__global__ void lk_emit(const int32_t* __restrict__ ea, const int32_t* __restrict__ eb, int64_t n_edges, int64_t n_nodes,
                        const int32_t* __restrict__ ra, const int32_t* __restrict__ pa,
                        const int32_t* __restrict__ rb, const int32_t* __restrict__ pb,
                        int group_bits, int pos_bits,
                        unsigned long long* counter, uint64_t* key1, uint64_t* key2, uint32_t* src) {
    const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n_edges + stride - 1) / stride;
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t e = e0 + it * stride;
        bool v1 = false, v2 = false;
        uint64_t k1a = 0, k2a = 0, k1b = 0, k2b = 0;
        if (e < n_edges) {
            const int32_t u = ea[e], v = eb[e];
            if (u >= 0 && v >= 0 && u < n_nodes && v < n_nodes) {
                const int32_t ru = ra[u], cv = rb[v];
                if (ru >= 0 && cv >= 0 && ru < cv) {
                    v1 = true;
                    k1a = ((uint64_t)(uint32_t)ru << group_bits) | (uint32_t)cv;
                    k2a = ((uint64_t)(uint32_t)pa[u] << pos_bits) | (uint32_t)pb[v];
                }
                if (u != v) {
                    const int32_t rv = ra[v], cu = rb[u];
                    if (rv >= 0 && cu >= 0 && rv < cu) {
                        v2 = true;
                        k1b = ((uint64_t)(uint32_t)rv << group_bits) | (uint32_t)cu;
                        k2b = ((uint64_t)(uint32_t)pa[v] << pos_bits) | (uint32_t)pb[u];
                    }
                }
            }
        }
        lk_append(v1, k1a, k2a, (uint32_t)e, counter, key1, key2, src);
        lk_append(v2, k1b, k2b, (uint32_t)e, counter, key1, key2, src);
    }
}

__global__ void lk_gather64(const uint64_t* __restrict__ in, const uint32_t* __restrict__ idx, int64_t m, uint64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = in[idx[i]];
}
__global__ void lk_iota(uint32_t* a, int64_t m) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) a[i] = (uint32_t)i;
}

// one thread per pair of groups: the reference's `weight += ...` chain in product order
__global__ void lk_sum(const uint64_t* __restrict__ ukeys, const int* __restrict__ run_len, const int* __restrict__ run_off, int n_runs,
                       const uint32_t* __restrict__ order2, const uint32_t* __restrict__ order1, const uint32_t* __restrict__ src,
                       const double* __restrict__ w, double cutoff, int group_bits,
                       int32_t* __restrict__ ga, int32_t* __restrict__ gb, double* __restrict__ total,
                       int64_t* __restrict__ n_edges, int64_t* __restrict__ n_over) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const int off = run_off[r], len = run_len[r];
    double sum = 0.0;
    int64_t over = 0;
    for (int i = 0; i < len; ++i) {
        // sorted slot -> slot after the first sort -> emitted item -> edge
        const uint32_t item = order1[order2[off + i]];
        sum = __dadd_rn(sum, w[src[item]]);
        over += sum > cutoff;
    }
    const uint64_t k = ukeys[r];
    ga[r] = (int32_t)(k >> group_bits);
    gb[r] = (int32_t)(k & ((1ull << group_bits) - 1));
    total[r] = sum; n_edges[r] = len; n_over[r] = over;
}

int bits_for(int64_t n) { int b = 1; while ((1LL << b) < n) ++b; return b; }

}  // namespace

void kb_links_free(kb_ctx* c) {
    cudaFree(c->d_lk_scratch);
    c->d_lk_scratch = nullptr; c->lk_scratch_bytes = 0;
    c->d_lk_a = c->d_lk_b = nullptr; c->d_lk_w = nullptr; c->d_lk_edges = c->d_lk_over = nullptr; c->lk_pairs = 0;
}

extern "C" int kb_links_build(kb_ctx* ctx, int64_t n_edges, const int32_t* d_a, const int32_t* d_b, const double* d_weight,
                              int64_t n_nodes, const int32_t* d_row_group, const int32_t* d_row_pos,
                              const int32_t* d_col_group, const int32_t* d_col_pos,
                              int64_t n_groups, int64_t max_pos, double cutoff, int64_t* n_pairs) {
    KB_CHECK_ARG(ctx && n_pairs, "null pointer");
    KB_CHECK_ARG(n_edges >= 0 && n_edges < (1LL << 30), "edge count (at most 2^30 edges)");
    KB_CHECK_ARG(n_nodes >= 0 && n_nodes < (1LL << 31) && n_groups >= 0 && n_groups < (1LL << 31) && max_pos >= 0 && max_pos < (1LL << 31), "sizes");
    KB_CHECK_ARG(n_edges == 0 || (d_a && d_b && d_weight), "edge arrays");
    KB_CHECK_ARG(n_nodes == 0 || (d_row_group && d_row_pos && d_col_group && d_col_pos), "role arrays");
    KB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->lk_pairs = 0;
    *n_pairs = 0;
    if (n_edges == 0 || n_nodes == 0 || n_groups < 2) return KB_OK;
    const int group_bits = bits_for(n_groups), pos_bits = bits_for(max_pos + 1);
    const int64_t cap = 2 * n_edges;                          // every edge is visited at most twice
    const int icap = (int)cap;
    // one grow-only scratch block per context: no allocation in the steady state
    size_t tb_sort = 0, tb_rle = 0, tb_scan = 0;
    KB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, icap, 0, 64, st));
    KB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tb_rle, (uint64_t*)nullptr, (uint64_t*)nullptr, (int*)nullptr, (int*)nullptr, icap, st));
    KB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb_scan, (int*)nullptr, (int*)nullptr, icap, st));
    size_t tb = tb_sort > tb_rle ? tb_sort : tb_rle;
    if (tb_scan > tb) tb = tb_scan;
    size_t off = 0;
    auto take = [&off](size_t bytes) { const size_t at = off; off += (size_t)kb_round_up((int64_t)bytes, 256); return at; };
    const size_t o_counter = take(8), o_nr = take(4);
    const size_t o_k1 = take((size_t)cap * 8), o_k2 = take((size_t)cap * 8), o_k2s = take((size_t)cap * 8);
    const size_t o_src = take((size_t)cap * 4), o_i0 = take((size_t)cap * 4), o_o1 = take((size_t)cap * 4), o_o2 = take((size_t)cap * 4);
    const size_t o_rl = take((size_t)cap * 4), o_ro = take((size_t)cap * 4);
    const size_t o_ga = take((size_t)cap * 4), o_gb = take((size_t)cap * 4);
    const size_t o_w = take((size_t)cap * 8), o_ne = take((size_t)cap * 8), o_ov = take((size_t)cap * 8);
    const size_t o_tmp = take(tb);
    if ((int64_t)off > ctx->lk_scratch_bytes) {
        KB_CUDA(cudaStreamSynchronize(st));
        cudaFree(ctx->d_lk_scratch); ctx->d_lk_scratch = nullptr; ctx->lk_scratch_bytes = 0;
        KB_CUDA(cudaMalloc(&ctx->d_lk_scratch, off));
        ctx->lk_scratch_bytes = (int64_t)off;
    }
    uint8_t* base = reinterpret_cast<uint8_t*>(ctx->d_lk_scratch);
    auto* counter = reinterpret_cast<unsigned long long*>(base + o_counter);
    int* nr = reinterpret_cast<int*>(base + o_nr);
    auto* k1 = reinterpret_cast<uint64_t*>(base + o_k1);      // (row group, column group) per item; later the unique pairs
    auto* k2 = reinterpret_cast<uint64_t*>(base + o_k2);      // (row position, column position); later k1 in sort-1 order
    auto* k2s = reinterpret_cast<uint64_t*>(base + o_k2s);    // sort outputs
    auto* src = reinterpret_cast<uint32_t*>(base + o_src);
    auto* i0 = reinterpret_cast<uint32_t*>(base + o_i0);
    auto* o1 = reinterpret_cast<uint32_t*>(base + o_o1);
    auto* o2 = reinterpret_cast<uint32_t*>(base + o_o2);
    int* rl = reinterpret_cast<int*>(base + o_rl);
    int* ro = reinterpret_cast<int*>(base + o_ro);
    void* tmp = base + o_tmp;
    ctx->d_lk_a = reinterpret_cast<int32_t*>(base + o_ga); ctx->d_lk_b = reinterpret_cast<int32_t*>(base + o_gb);
    ctx->d_lk_w = reinterpret_cast<double*>(base + o_w);
    ctx->d_lk_edges = reinterpret_cast<int64_t*>(base + o_ne); ctx->d_lk_over = reinterpret_cast<int64_t*>(base + o_ov);

    KbTimer timer(ctx, 8);
    KB_CUDA(cudaMemsetAsync(counter, 0, 8, st));
    int64_t blocks = (n_edges + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    lk_emit<<<(unsigned)blocks, 256, 0, st>>>(d_a, d_b, n_edges, n_nodes, d_row_group, d_row_pos, d_col_group, d_col_pos,
                                              group_bits, pos_bits, counter, k1, k2, src);
    ctx->launches++;
    KB_CUDA(cudaGetLastError());
    unsigned long long h_m = 0;
    KB_CUDA(cudaMemcpyAsync(&h_m, counter, 8, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (h_m == 0) return KB_OK;
    const int m = (int)h_m;
    const unsigned gm = (unsigned)((m + 255) / 256);
    // sort 1: by (row position, column position); payload = item number
    lk_iota<<<gm, 256, 0, st>>>(i0, m);
    KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k2, k2s, i0, o1, m, 0, 2 * pos_bits, st));
    // sort 2 (stable): by (row group, column group); payload = slot of sort 1
    lk_gather64<<<gm, 256, 0, st>>>(k1, o1, m, k2);
    KB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k2, k2s, i0, o2, m, 0, 2 * group_bits, st));
    // runs of equal (row group, column group) and where they start
    KB_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tb, k2s, k1, rl, nr, m, st));
    ctx->launches += 5;
    int h_r = 0;
    KB_CUDA(cudaMemcpyAsync(&h_r, nr, 4, cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (h_r <= 0) { kb_set_error("internal: no runs for %d items", m); return KB_ECUDA; }
    KB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, rl, ro, h_r, st));
    lk_sum<<<(unsigned)((h_r + 127) / 128), 128, 0, st>>>(k1, rl, ro, h_r, o2, o1, src, d_weight, cutoff, group_bits,
                                                          ctx->d_lk_a, ctx->d_lk_b, ctx->d_lk_w, ctx->d_lk_edges, ctx->d_lk_over);
    ctx->launches += 2;
    KB_CUDA(cudaGetLastError());
    KB_CUDA(cudaStreamSynchronize(st));
    ctx->lk_pairs = h_r;
    *n_pairs = h_r;
    return KB_OK;
}

extern "C" int kb_links_fetch(kb_ctx* ctx, int32_t* h_group_a, int32_t* h_group_b, double* h_weight, int64_t* h_edges, int64_t* h_over) {
    KB_CHECK_ARG(ctx, "ctx");
    const size_t r = (size_t)ctx->lk_pairs;
    if (r == 0) return KB_OK;
    KB_CUDA(cudaSetDevice(ctx->device));
    if (h_group_a) KB_CUDA(cudaMemcpy(h_group_a, ctx->d_lk_a, r * 4, cudaMemcpyDeviceToHost));
    if (h_group_b) KB_CUDA(cudaMemcpy(h_group_b, ctx->d_lk_b, r * 4, cudaMemcpyDeviceToHost));
    if (h_weight) KB_CUDA(cudaMemcpy(h_weight, ctx->d_lk_w, r * 8, cudaMemcpyDeviceToHost));
    if (h_edges) KB_CUDA(cudaMemcpy(h_edges, ctx->d_lk_edges, r * 8, cudaMemcpyDeviceToHost));
    if (h_over) KB_CUDA(cudaMemcpy(h_over, ctx->d_lk_over, r * 8, cudaMemcpyDeviceToHost));
    return KB_OK;
}
