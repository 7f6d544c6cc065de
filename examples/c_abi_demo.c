/* Plain-C host program against include/karma_b200.h: what a non-Python host (the binding a maintainer
 * would write in any FFI) does for the hot path.  Device buffers come from the CUDA runtime here; in
 * karma they come from torch.  Build:
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_demo.c \
 *       -L karma_b200 -lkarma_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/karma_b200 -o c_abi_demo
 * Runs on a B200 only (kb_create fails loudly elsewhere: there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "karma_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, kb_last_error()); return 1; } } while (0)
#define CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(void) {
    /* the dict of karma.py:40-61, already packed: two contigs, header keys ">c1" and ">contig_two" */
    const char* seqs[2] = {"ACGTACGTTGCAACGTACGT", "AAAAAAAAAATTTTTTTTTT"};
    const int32_t key_len[2] = {3, 11};
    int64_t offsets[3] = {0, 0, 0};
    char bases[128];
    memset(bases, 0, sizeof bases);
    for (int i = 0; i < 2; ++i) { memcpy(bases + offsets[i], seqs[i], strlen(seqs[i])); offsets[i + 1] = offsets[i] + (int64_t)strlen(seqs[i]); }

    kb_ctx* ctx = NULL;
    CHECK(kb_create(&ctx, 0));
    const int32_t D = kb_mode_columns(KB_MODE_5P6);              /* 1088 */
    const int64_t N = 2;
    uint8_t* d_bases; int64_t* d_off; int32_t* d_len; uint32_t *d_counts, *d_exotic, *d_presence; double* d_profile;
    CUDA(cudaMalloc((void**)&d_bases, sizeof bases)); CUDA(cudaMalloc((void**)&d_off, sizeof offsets)); CUDA(cudaMalloc((void**)&d_len, sizeof key_len));
    CUDA(cudaMalloc((void**)&d_counts, (size_t)N * D * 4)); CUDA(cudaMalloc((void**)&d_exotic, (size_t)N * 4));
    CUDA(cudaMalloc((void**)&d_presence, (size_t)(D + 1) * 4)); CUDA(cudaMalloc((void**)&d_profile, (size_t)N * D * 8));
    CUDA(cudaMemcpy(d_bases, bases, sizeof bases, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(d_off, offsets, sizeof offsets, cudaMemcpyHostToDevice));
    CUDA(cudaMemcpy(d_len, key_len, sizeof key_len, cudaMemcpyHostToDevice));
    CUDA(cudaMemset(d_presence, 0, (size_t)(D + 1) * 4));

    CHECK(kb_count(ctx, KB_MODE_5P6, d_bases, d_off, N, d_counts, D, d_exotic, d_presence));          /* kmer.py:56-92 */
    CHECK(kb_normalise(ctx, d_counts, D, D, d_len, N, N, d_profile, D, NULL, 0, NULL, NULL, NULL));                  /* count / len(key) */
    double* profile = (double*)malloc((size_t)N * D * 8);
    CUDA(cudaMemcpy(profile, d_profile, (size_t)N * D * 8, cudaMemcpyDeviceToHost));
    /* column 0 = "AAAAA", column 1 = "AAAAAA" (sorted order of kmer.py:172): 6/11 and 5/11 for >contig_two */
    printf("AAAAA  of >contig_two: %.17g (expect %.17g)\n", profile[D + 0], 6.0 / 11.0);
    printf("AAAAAA of >contig_two: %.17g (expect %.17g)\n", profile[D + 1], 5.0 / 11.0);
    const int ok = profile[D + 0] == 6.0 / 11.0 && profile[D + 1] == 5.0 / 11.0;
    free(profile);
    cudaFree(d_bases); cudaFree(d_off); cudaFree(d_len); cudaFree(d_counts); cudaFree(d_exotic); cudaFree(d_presence); cudaFree(d_profile);
    CHECK(kb_destroy(ctx));
    printf(ok ? "c_abi_demo OK\n" : "c_abi_demo MISMATCH\n");
    return ok ? 0 : 2;
}
