// Parameter block shared by the two tcgen05 GEMM kernels (gemm.cu: one tile per CTA, fully general; gemm2.cu:
// persistent, double-buffered TMEM accumulators, coalesced epilogue).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace mobi {

struct GemmParams {
    int M, N;
    int num_k_blocks;
    // conv
    int conv;
    int C, H, W, KW, pad_h, pad_w, cblocks;
    // epilogue
    void* out;
    void* out2;
    void* out3;
    const float* bias;
    const float* row_bias;
    const void* residual;
    long long ldo;
    long long ld_row_bias;
    int rows_per_group;
    int out_f32, res_f32;
    int mode;
    int act;
    int heads, head_dim, tokens;
    long long out_seg, out_seg_stride, out_seg_offset;
    int res_prefetch;             // 1: the producer pulls each tile's residual box into L2 ahead of the epilogue (tmR is valid)
    int pair;                     // 1: run as 2-CTA clusters (cta_group::2, 256-row tiles, B box = BN / 2 rows); 2: two such pairs
                                  // per cluster on neighbouring n-tiles, A tiles multicast between them (A box = 64 rows)
    int batch;                    // > 1: batched problem, A/B through 3-D tensor maps
    long long out_batch_stride;   // elements
    int batch_inner;              // > 0: two-level batch, z -> (z % batch_inner, z / batch_inner) over 4-D operand maps
    long long out_batch2_stride;  // elements between outer batch entries of the output
    int atomic_out;               // f32 output accumulated with atomic adds (split-K over the batch index)
    int a_mn, b_mn;               // operand given MN-major ([K, M] / [K, N] row-major): 64 x 64 TMA boxes, UMMA major bits
    float* colstats;              // optional f32 [2][ceil(M / 32)][N]: column sums / sums of squares per 32-row group
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;

// true when the persistent kernel's vectorised epilogue can take this problem
bool gemm2_supported(const GemmParams& p);
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p, int bn_tile,
                 cudaStream_t stream);
// true when the 2-CTA (cta_group::2) variant should run this problem with tile width bn_tile
bool gemm2_pair_wanted(const GemmParams& p, int bn_tile, int pair_request);
bool gemm2_quad_ok(const GemmParams& p, int bn_tile);

}  // namespace mobi
