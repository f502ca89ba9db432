// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[M,N] = A[M,K] . B[N,K]^T     bf16 operands, fp32 accumulation in TMEM
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: streams 128x64 A tiles and BNx64 B tiles (128B-swizzled) through a
//               STAGES-deep shared-memory ring guarded by full/empty mbarriers;
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage and
//               releases the stage with tcgen05.commit; also owns the TMEM allocation;
//   warps 2-5   epilogue: tcgen05.ld the accumulator (lane = output row), apply bias / row-group bias /
//               residual / GEGLU, and store in one of the layouts the attention kernels want.
// With `conv` set, the A tile of filter tap (kh,kw) is a shifted 4-D TMA box over the NHWC image; the
// out-of-bounds halo is zero-filled by the TMA unit, so no im2col buffer exists.
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace mobi {

template <int BN>
struct GemmCfg {
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGES = (BN <= 64) ? 4 : (BN <= 128 ? 6 : 4);
    static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    static constexpr int SMEM_BYTES = STAGES * (A_TILE_BYTES + B_TILE_BYTES) + 256 + 1024;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Stores 8 consecutive output columns [n0, n0+8) of row m.  v already has bias etc. applied.
__device__ __forceinline__ void store8(const GemmParams& p, long long m, int n0, const float (&v)[8]) {
    if (p.mode == MOBI_EPI_PLAIN) {
        const long long off = m * p.ldo + n0;  // m is already the (remapped) output row
        const bool full = (n0 + 8 <= p.N);
        if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + off;
            if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
                reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (n0 + j < p.N) o[j] = v[j];
            }
        } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
            if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                uint4 pk;
                pk.x = pack_bf16x2(v[0], v[1]);
                pk.y = pack_bf16x2(v[2], v[3]);
                pk.z = pack_bf16x2(v[4], v[5]);
                pk.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(o) = pk;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (n0 + j < p.N) o[j] = __float2bfloat16(v[j]);
            }
        }
        return;
    }
    // head-split layouts (n0 is a multiple of 8 and head_dim % 8 == 0, so the 8 columns share a head)
    if (n0 >= p.N) return;
    const int inner = p.heads * p.head_dim;
    int which = 0, n = n0;
    int mode = p.mode;
    if (mode == MOBI_EPI_QKV || mode == MOBI_EPI_KV) {
        which = n0 / inner;
        n = n0 - which * inner;
        const int last = (mode == MOBI_EPI_QKV) ? 2 : 1;
        mode = (which == last) ? MOBI_EPI_HEADS_T : MOBI_EPI_HEADS;
    } else if (mode == MOBI_EPI_QKV_ROW || mode == MOBI_EPI_KV_ROW) {
        which = n0 / inner;
        n = n0 - which * inner;
        mode = MOBI_EPI_HEADS;
    }
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(which == 0 ? p.out : (which == 1 ? p.out2 : p.out3));
    const int h = n / p.head_dim;
    const int dd = n - h * p.head_dim;
    const long long b = m / p.tokens;
    const long long t = m - b * p.tokens;
    const long long bh = b * p.heads + h;
    if (mode == MOBI_EPI_HEADS) {
        __nv_bfloat16* o = base + (bh * p.tokens + t) * p.head_dim + dd;
        uint4 pk;
        pk.x = pack_bf16x2(v[0], v[1]);
        pk.y = pack_bf16x2(v[2], v[3]);
        pk.z = pack_bf16x2(v[4], v[5]);
        pk.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(o) = pk;  // (head_dim % 8 == 0) keeps this 16-byte aligned
    } else {
        __nv_bfloat16* o = base + (bh * p.head_dim + dd) * p.tokens + t;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[(long long)j * p.tokens] = __float2bfloat16(v[j]);
    }
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_TILE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_TILE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    // m tiles on grid.x (2^31 - 1 limit): the VAE decode at 512^2 has > 65535 of them per batch
    const int m_tile = blockIdx.x;
    const int n_tile = blockIdx.y;

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_init(accum_bar, 1);
            fence_barrier_init();
        }
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = p.num_k_blocks;

    if (warp == 0) {
        if (elect_one()) {
            // ---------------- TMA producer
            int x0 = 0, y0 = 0, n0 = 0;
            if (p.conv) {
                const long long pix = (long long)m_tile * BM;
                x0 = (int)(pix % p.W);
                y0 = (int)((pix / p.W) % p.H);
                n0 = (int)(pix / ((long long)p.W * p.H));
            }
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], A_TILE_BYTES + Cfg::B_TILE_BYTES);
                if (p.conv) {
                    const int tap = kb / p.cblocks;
                    const int cb = kb - tap * p.cblocks;
                    const int kh = tap / p.KW;
                    const int kw = tap - kh * p.KW;
                    tma_load_4d(sA + s * A_TILE_BYTES, &tmA, &full_bar[s], cb * BK, x0 + kw - p.pad_w,
                                y0 + kh - p.pad_h, n0);
                    tma_load_2d(sB + s * Cfg::B_TILE_BYTES, &tmB, &full_bar[s], tap * p.C + cb * BK, n_tile * BN);
                } else {
                    tma_load_2d(sA + s * A_TILE_BYTES, &tmA, &full_bar[s], kb * BK, m_tile * BM);
                    tma_load_2d(sB + s * Cfg::B_TILE_BYTES, &tmB, &full_bar[s], kb * BK, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ---------------- MMA issuer
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA + s * A_TILE_BYTES));
                const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sB + s * Cfg::B_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // +32 bytes along K inside the 128-byte swizzle row = +2 in the (addr >> 4) field
                    umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        // ---------------- epilogue (4 warps; warp w may only touch TMEM lanes [32*(w%4), +32))
        const int lg = warp & 3;
        const long long m = (long long)m_tile * BM + lg * 32 + lane;
        const bool row_ok = m < p.M;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const float* rb = (p.row_bias && row_ok) ? p.row_bias + (m / p.rows_per_group) * p.ld_row_bias : nullptr;
        // output row: identity, or gathered segments (camera-only / lidar-only rows of the interleaved batch)
        const long long mo = (p.out_seg > 0 && (p.mode <= MOBI_EPI_GEGLU || p.mode == MOBI_EPI_GEGLU2))
                                 ? (m / p.out_seg) * p.out_seg_stride + p.out_seg_offset + (m % p.out_seg)
                                 : m;
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            uint32_t r[16];
            __syncwarp();
            tmem_ld16(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + c, r);
            tmem_ld_wait();
            const int n0 = n_tile * BN + c;
            if (!row_ok || n0 >= p.N) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
            }
            if (rb) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (n0 + j < p.N) v[j] += __ldg(rb + n0 + j);
            }
            if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = v[j] / (1.0f + __expf(-v[j]));
            } else if (p.act == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
            }
            if (p.mode == MOBI_EPI_GEGLU2) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = v[2 * j] * gelu_erf(v[2 * j + 1]);
                GemmParams q = p;
                q.mode = MOBI_EPI_PLAIN;
                q.N = p.N / 2;
                store8(q, mo, n0 / 2, o);
                continue;
            }
            if (p.mode == MOBI_EPI_GEGLU) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = v[j] * gelu_erf(v[8 + j]);
                // output has N/2 columns
                GemmParams q = p;
                q.mode = MOBI_EPI_PLAIN;
                q.N = p.N / 2;
                store8(q, mo, n0 / 2, o);
                continue;
            }
            if (p.residual) {
                const long long off = mo * p.ldo + n0;
                if (p.res_f32) {
                    const float* rs = reinterpret_cast<const float*>(p.residual) + off;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < p.N) v[j] += rs[j];
                } else {
                    const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(p.residual) + off;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n0 + j < p.N) v[j] += __bfloat162float(rs[j]);
                }
            }
            float lo[8], hi[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                lo[j] = v[j];
                hi[j] = v[8 + j];
            }
            store8(p, mo, n0, lo);
            store8(p, mo, n0 + 8, hi);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
        configured = true;
    }
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, 1);
    gemm_bf16_kernel<BN><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mobi

using namespace mobi;

// Operand tensor map for the (possibly batched, possibly two-level batched) GEMM: dim0 x dim1 matrix with row stride ld,
// box0 x box1 tiles; batch entries `bs` elements apart, outer batch entries `bs2` elements apart.
static int make_operand_map(CUtensorMap* tm, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box0,
                            uint32_t box1, int batch, int inner, int64_t bs, int64_t bs2) {
    if (batch <= 1) {
        uint64_t dims[2] = {dim0, dim1};
        uint64_t strides[1] = {ld * 2};
        uint32_t box[2] = {box0, box1};
        return make_tensor_map_bf16(tm, base, 2, dims, strides, box);
    }
    if (inner <= 0) {
        uint64_t dims[3] = {dim0, dim1, (uint64_t)batch};
        uint64_t strides[2] = {ld * 2, (uint64_t)bs * 2};
        uint32_t box[3] = {box0, box1, 1};
        return make_tensor_map_bf16(tm, base, 3, dims, strides, box);
    }
    uint64_t dims[4] = {dim0, dim1, (uint64_t)inner, (uint64_t)(batch / inner)};
    uint64_t strides[3] = {ld * 2, (uint64_t)bs * 2, (uint64_t)bs2 * 2};
    uint32_t box[4] = {box0, box1, 1, 1};
    return make_tensor_map_bf16(tm, base, 4, dims, strides, box);
}

// `plan` != nullptr: validate the problem and decide tile shape / CTA grouping exactly as a launch would, but build no
// tensor maps and launch nothing (no device needed): plan[0] = tile_n, plan[1] = pair mode (0 / 1 / 2 / 3), plan[2] = 1 when
// the persistent kernel takes the problem (0: the one-tile kernel).
static int gemm_impl(const mobi_gemm_args* a, cudaStream_t stream, int32_t* plan) {
    MOBI_CHECK(a != nullptr, "mobi_gemm: null args");
    MOBI_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "mobi_gemm: empty problem M=%lld N=%lld K=%lld", (long long)a->M,
               (long long)a->N, (long long)a->K);
    MOBI_CHECK(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), "mobi_gemm: dims must fit int32");
    MOBI_CHECK(a->A && a->B && a->out, "mobi_gemm: null operand");
    GemmParams p{};
    p.M = (int)a->M;
    p.N = (int)a->N;
    p.out = a->out;
    p.out2 = a->out2;
    p.out3 = a->out3;
    p.bias = a->bias;
    p.row_bias = a->row_bias;
    p.residual = a->residual;
    p.ldo = a->ldo;
    p.rows_per_group = a->rows_per_group > 0 ? (int)a->rows_per_group : 1;
    p.ld_row_bias = a->ld_row_bias > 0 ? a->ld_row_bias : a->N;
    p.act = a->act;
    MOBI_CHECK(a->act >= 0 && a->act <= 2, "mobi_gemm: act=%d (0 none, 1 SiLU, 2 GELU)", a->act);
    p.out_f32 = a->out_dtype == MOBI_DTYPE_F32;
    p.res_f32 = a->res_dtype == MOBI_DTYPE_F32;
    p.mode = a->epilogue;
    p.heads = a->heads;
    p.head_dim = a->head_dim;
    p.tokens = a->tokens;
    p.out_seg = a->out_seg;
    p.out_seg_stride = a->out_seg_stride;
    p.out_seg_offset = a->out_seg_offset;
    MOBI_CHECK(p.mode >= MOBI_EPI_PLAIN && p.mode <= MOBI_EPI_KV_ROW, "mobi_gemm: bad epilogue %d", p.mode);
    const bool batched = a->batch > 1;
    p.batch = batched ? a->batch : 1;
    p.out_batch_stride = batched ? a->out_batch_stride : 0;
    p.batch_inner = (batched && a->batch_inner > 0) ? a->batch_inner : 0;
    p.out_batch2_stride = p.batch_inner ? a->out_batch2_stride : 0;
    if (p.batch_inner)
        MOBI_CHECK(a->batch % a->batch_inner == 0 && a->a_batch2_stride % 8 == 0 && a->b_batch2_stride % 8 == 0 &&
                       a->out_batch2_stride % 4 == 0 && a->kernel != 1,
                   "mobi_gemm: batch_inner must divide batch; outer batch strides keep 16-byte alignment");
    if (batched) {
        MOBI_CHECK(p.mode == MOBI_EPI_PLAIN && !a->conv && a->out_seg == 0 && a->row_bias == nullptr,
                   "mobi_gemm: batch needs the PLAIN epilogue without conv / row segments / row_bias");
        MOBI_CHECK(a->a_batch_stride % 8 == 0 && a->b_batch_stride % 8 == 0 && a->out_batch_stride % 4 == 0,
                   "mobi_gemm: batch strides must keep 16-byte alignment (A, B: %% 8, out: %% 4 elements)");
    }
    if (p.mode == MOBI_EPI_GEGLU) {
        MOBI_CHECK(a->N % 16 == 0, "mobi_gemm: GEGLU needs N %% 16 == 0 (N=%lld)", (long long)a->N);
        MOBI_CHECK(a->residual == nullptr, "mobi_gemm: GEGLU epilogue takes no residual");
    }
    if (p.mode == MOBI_EPI_GEGLU2) {
        MOBI_CHECK(a->N % 16 == 0, "mobi_gemm: GEGLU2 needs N %% 16 == 0 (N=%lld)", (long long)a->N);
        MOBI_CHECK(a->residual == nullptr, "mobi_gemm: GEGLU epilogue takes no residual");
    }
    if ((p.mode >= MOBI_EPI_HEADS && p.mode <= MOBI_EPI_KV) || p.mode >= MOBI_EPI_QKV_ROW) {
        MOBI_CHECK(p.heads > 0 && p.head_dim > 0 && p.tokens > 0 && p.head_dim % 8 == 0,
                   "mobi_gemm: head layouts need heads, head_dim %% 8 == 0, tokens");
        const long long inner = (long long)p.heads * p.head_dim;
        const long long parts = (p.mode == MOBI_EPI_QKV || p.mode == MOBI_EPI_QKV_ROW)
                                    ? 3
                                    : ((p.mode == MOBI_EPI_KV || p.mode == MOBI_EPI_KV_ROW) ? 2 : 1);
        MOBI_CHECK(a->N == parts * inner, "mobi_gemm: N=%lld does not match heads*d", (long long)a->N);
        MOBI_CHECK(a->M % p.tokens == 0, "mobi_gemm: M must be a multiple of tokens");
        MOBI_CHECK(a->out_dtype == MOBI_DTYPE_BF16 && a->residual == nullptr, "mobi_gemm: head layouts are bf16");
        if (p.mode == MOBI_EPI_QKV || p.mode == MOBI_EPI_QKV_ROW) MOBI_CHECK(a->out2 && a->out3, "mobi_gemm: QKV needs out2/out3");
        if (p.mode == MOBI_EPI_KV || p.mode == MOBI_EPI_KV_ROW) MOBI_CHECK(a->out2 != nullptr, "mobi_gemm: KV needs out2");
    }

    CUtensorMap tmA, tmB;
    int64_t K = a->K;
    int bw = 0, bh = 0, bn = 0;   // conv: the 128 pixels of an A tile as a (w, h, image) box
    if (a->conv) {
        MOBI_CHECK(a->C % 8 == 0, "mobi_gemm(conv): C=%d must be a multiple of 8", a->C);
        MOBI_CHECK(a->KH >= 1 && a->KW >= 1 && (long long)a->n_img * a->H * a->W == a->M,
                   "mobi_gemm(conv): M != n_img*H*W");
        K = (int64_t)a->KH * a->KW * a->C;
        MOBI_CHECK(a->K == K, "mobi_gemm(conv): K=%lld != KH*KW*C=%lld", (long long)a->K, (long long)K);
        if (a->W >= BM) {
            MOBI_CHECK(a->W % BM == 0, "mobi_gemm(conv): W=%d must be a multiple of 128 (or <= 128 power of 2)", a->W);
            bw = BM;
            bh = 1;
            bn = 1;
        } else {
            MOBI_CHECK(BM % a->W == 0, "mobi_gemm(conv): W=%d must divide 128", a->W);
            bw = a->W;
            bh = BM / bw;
            if (bh > a->H) bh = a->H;
            MOBI_CHECK(a->H % bh == 0 && BM % (bw * bh) == 0, "mobi_gemm(conv): H=%d W=%d not tileable", a->H, a->W);
            bn = BM / (bw * bh);
        }
        p.conv = 1;
        p.C = a->C;
        p.H = a->H;
        p.W = a->W;
        p.KW = a->KW;
        p.pad_h = a->pad_h;
        p.pad_w = a->pad_w;
        p.cblocks = (a->C + BK - 1) / BK;
        p.num_k_blocks = a->KH * a->KW * p.cblocks;
        uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
        uint64_t strides[3] = {(uint64_t)a->C * 2, (uint64_t)a->W * a->C * 2, (uint64_t)a->H * a->W * a->C * 2};
        uint32_t box[4] = {BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
        if (!plan && make_tensor_map_bf16(&tmA, a->A, 4, dims, strides, box)) return 1;
    } else if (a->a_mn_major) {
        MOBI_CHECK(a->lda % 8 == 0 && a->lda >= a->M, "mobi_gemm: MN-major A needs lda=%lld >= M and a multiple of 8",
                   (long long)a->lda);
        p.num_k_blocks = (int)((K + BK - 1) / BK);
        if (!plan && make_operand_map(&tmA, a->A, (uint64_t)a->M, (uint64_t)K, (uint64_t)a->lda, 64, BK, p.batch, p.batch_inner,
                                      a->a_batch_stride, a->a_batch2_stride))
            return 1;
    } else {
        MOBI_CHECK(a->K % 8 == 0 && a->lda % 8 == 0, "mobi_gemm: K=%lld and lda=%lld must be multiples of 8",
                   (long long)a->K, (long long)a->lda);
        p.num_k_blocks = (int)((K + BK - 1) / BK);
        if (!plan && make_operand_map(&tmA, a->A, (uint64_t)K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM, p.batch, p.batch_inner,
                                      a->a_batch_stride, a->a_batch2_stride))
            return 1;
    }
    p.atomic_out = a->atomic_out ? 1 : 0;
    p.colstats = a->colstats;
    if (p.colstats)
        MOBI_CHECK(p.mode == MOBI_EPI_PLAIN && p.out_f32 && !batched && a->out_seg == 0 && !a->atomic_out && a->kernel != 1 &&
                       a->N % 4 == 0,
                   "mobi_gemm: colstats needs the PLAIN epilogue with f32 output on the persistent kernel (no batch, row "
                   "segments or atomic accumulation)");
    if (p.atomic_out)
        MOBI_CHECK(p.out_f32 && p.mode == MOBI_EPI_PLAIN && !a->bias && !a->row_bias && !a->residual && a->act == 0 &&
                       a->kernel != 1 && a->out_seg == 0,
                   "mobi_gemm: atomic_out needs a plain f32 output without bias / activation / residual");
    p.a_mn = a->a_mn_major ? 1 : 0;
    p.b_mn = a->b_mn_major ? 1 : 0;
    if (p.a_mn || p.b_mn) {
        MOBI_CHECK(!a->conv && a->kernel != 1 && a->pair != 1, "mobi_gemm: MN-major operands need the persistent kernel, no conv, no CTA pairs");
        MOBI_CHECK(a->tile_n == 0 || !p.b_mn || a->tile_n % 64 == 0, "mobi_gemm: MN-major B needs tile_n %% 64 == 0");
    }
    if (p.b_mn) MOBI_CHECK(a->ldb % 8 == 0 && a->ldb >= a->N, "mobi_gemm: MN-major B needs ldb=%lld >= N and a multiple of 8", (long long)a->ldb);
    else MOBI_CHECK(a->ldb % 8 == 0 && a->ldb >= K, "mobi_gemm: ldb=%lld must be a multiple of 8 and >= K", (long long)a->ldb);

    int bn_tile = a->tile_n;
    if (bn_tile == 0) {
        // Largest tile that keeps waste low and still gives every SM at least one CTA.
        const long long mt = (a->M + BM - 1) / BM * p.batch;
        const int cand[4] = {256, 160, 128, 64};
        bn_tile = 64;
        double best = 1e30;
        for (int i = 0; i < 4; ++i) {
            const int bnc = cand[i];
            if (p.b_mn && bnc % 64 != 0) continue;  // MN-major B tiles are made of 64-column boxes
            const long long nt = (a->N + bnc - 1) / bnc;
            const long long ctas = mt * nt;
            const long long waves = (ctas + sm_count() - 1) / sm_count();
            // cost model: waves * (tile work) ; small tiles pay an shared-memory-bandwidth penalty
            const double eff = bnc >= 160 ? 1.0 : (bnc == 128 ? 0.9 : 0.67);
            const double cost = (double)waves * bnc / eff;
            if (cost < best - 1e-9) {
                best = cost;
                bn_tile = bnc;
            }
        }
    }
    int pair_request = a->pair;
    if (a->tile_n == 0 && a->pair == 0 && a->kernel != 1 && !a->conv && p.mode != MOBI_EPI_PLAIN && !p.a_mn && !p.b_mn &&
        gemm2_supported(p)) {
        // bf16-output epilogues (head-split QKV, GEGLU): with K = 320..1280 these are bound by the L2 -> shared-memory
        // traffic of the operand tiles and by the epilogue, not by the MMAs.  A CTA pair on a 256 x 256 tile moves 32 KB
        // per 128 x 256 x 64 MMA block instead of 36 KB per 128 x 160 x 64 (0.55x the bytes per FLOP); measured on the
        // UNet's shapes it is 4-16 % faster wherever N fills the wide tiles (profiles/r02/kbench_gemm_tiles.log).
        const long long nt = (a->N + 255) / 256, m2 = (a->M + 2 * BM - 1) / (2 * BM);
        if (nt * 256 * 10 <= (long long)a->N * 11 && m2 * nt * p.batch >= sm_count() / 2) {
            bn_tile = 256;
            pair_request = 1;
        }
    }
    if (a->tile_n == 0 && a->pair == 0 && a->kernel != 1 && a->conv && a->N % 320 == 0 && p.num_k_blocks >= 16 &&
        p.batch <= 1 && gemm2_supported(p)) {
        // 3x3 convolutions whose width is a multiple of 320 (every ResBlock of the UNet): a CTA pair takes 256 x 320 tiles as
        // two N = 160 MMAs per k-step on ONE A tile, which cuts the operand bytes delivered to an SM per FLOP by 31 % --
        // what bounds these kernels (profiles/r02/ncu_gemm_l2_bound.md).  1.15-1.5x on the UNet's shapes once there is a
        // full wave of such tiles (profiles/r02/kbench_gemm_tiles.log).
        const long long m2 = (a->M + 2 * BM - 1) / (2 * BM);
        if (m2 * (a->N / 320) >= sm_count() / 2) {
            bn_tile = 160;
            pair_request = 3;
        }
    }
    p.pair = (a->kernel != 1 && gemm2_supported(p) && gemm2_pair_wanted(p, bn_tile, pair_request)) ? 1 : 0;
    MOBI_CHECK(a->pair < 1 || p.pair, "mobi_gemm: pair = 1 / 2 / 3 needs a problem the persistent kernel supports");
    if (a->pair == 2) {
        MOBI_CHECK(gemm2_quad_ok(p, bn_tile), "mobi_gemm: pair = 2 (4-CTA clusters) needs the PLAIN epilogue, tile_n >= 128, an "
                                              "even number of n-tiles, K-major operands and no batch");
        p.pair = 2;
    }
    if (a->pair == 3 || pair_request == 3) {
        MOBI_CHECK(p.pair && bn_tile == 160 && p.batch <= 1 && !p.a_mn && !p.b_mn,
                   "mobi_gemm: pair = 3 (320-column tiles on CTA pairs) needs tile_n = 160, K-major operands and no batch");
        p.pair = 3;
    }
    if (plan) {
        plan[0] = bn_tile;
        plan[1] = p.pair;
        plan[2] = (a->kernel != 1 && gemm2_supported(p)) ? 1 : 0;
        return 0;
    }
    if (p.pair == 2) {
        // the A box shrinks to the 64 rows each CTA fetches (and multicasts to its counterpart)
        if (a->conv) {
            uint32_t hw = (uint32_t)bw, hh = (uint32_t)bh, hn = (uint32_t)bn;
            if (hn >= 2) hn /= 2;
            else if (hh >= 2) hh /= 2;
            else hw /= 2;
            uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
            uint64_t strides[3] = {(uint64_t)a->C * 2, (uint64_t)a->W * a->C * 2, (uint64_t)a->H * a->W * a->C * 2};
            uint32_t box[4] = {BK, hw, hh, hn};
            if (make_tensor_map_bf16(&tmA, a->A, 4, dims, strides, box)) return 1;
        } else {
            if (make_operand_map(&tmA, a->A, (uint64_t)K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM / 2, 1, 0, 0, 0)) return 1;
        }
    }
    if (p.b_mn) {
        if (make_operand_map(&tmB, a->B, (uint64_t)a->N, (uint64_t)K, (uint64_t)a->ldb, 64, BK, p.batch, p.batch_inner,
                             a->b_batch_stride, a->b_batch2_stride))
            return 1;
    } else {
        // a CTA pair splits the B tile
        if (make_operand_map(&tmB, a->B, (uint64_t)K, (uint64_t)a->N, (uint64_t)a->ldb, BK, (uint32_t)(p.pair ? bn_tile / 2 : bn_tile),
                             p.batch, p.batch_inner, a->b_batch_stride, a->b_batch2_stride))
            return 1;
    }
    MOBI_CHECK(!(p.a_mn || p.b_mn || p.atomic_out) || gemm2_supported(p),
               "mobi_gemm: MN-major operands / atomic_out are outside the persistent kernel's epilogue here");
    if (a->kernel != 1 && gemm2_supported(p)) {
        // residual boxes go to L2 ahead of the epilogue (one TMA prefetch per tile by the producer thread)
        CUtensorMap tmR = tmA;
        const int rb = p.res_f32 ? 4 : 2;
        const uint32_t box_cols = (uint32_t)(bn_tile < p.N ? bn_tile : p.N);
        // rows of the output / residual buffer the mapped tiles can touch (row segments spread them out)
        const long long out_rows = p.out_seg > 0 ? ((long long)(p.M + p.out_seg - 1) / p.out_seg - 1) * p.out_seg_stride +
                                                       p.out_seg_offset + p.out_seg
                                                 : (long long)p.M;
        if (p.residual && p.mode == MOBI_EPI_PLAIN && (p.out_seg % BM) == 0 && out_rows < (1ll << 31) && p.batch <= 1 && !p.atomic_out &&
            ((long long)p.ldo * rb) % 16 == 0 && (box_cols * rb) % 16 == 0 && p.num_k_blocks <= 10 && res_prefetch_enabled()) {
            // short K only (measured: to_out + residual, K = 320 / 640: 4 % faster; with K >= 1280 the operand stream
            // already fills HBM and the extra early traffic costs 3-5 %)
            if (make_tensor_map_2d_plain(&tmR, p.residual, rb, (uint64_t)p.N, (uint64_t)out_rows, (uint64_t)p.ldo * rb, box_cols, BM))
                return 1;
            p.res_prefetch = 1;
        }
        return launch_gemm2(tmA, tmB, tmR, p, bn_tile, stream);
    }
    MOBI_CHECK(p.colstats == nullptr, "mobi_gemm: colstats requested but this problem falls back to the one-tile kernel");
    MOBI_CHECK(!batched, "mobi_gemm: this batched problem is outside the persistent kernel's epilogue (N %% 4, "
                         "16-byte aligned out / residual / bias)");
    switch (bn_tile) {
        case 64: return launch_gemm<64>(tmA, tmB, p, stream);
        case 128: return launch_gemm<128>(tmA, tmB, p, stream);
        case 160: return launch_gemm<160>(tmA, tmB, p, stream);
        case 256: return launch_gemm<256>(tmA, tmB, p, stream);
        default: MOBI_CHECK(false, "mobi_gemm: unsupported tile_n %d", bn_tile);
    }
    return 0;
}

extern "C" int mobi_gemm(const mobi_gemm_args* a, void* stream_) {
    return gemm_impl(a, reinterpret_cast<cudaStream_t>(stream_), nullptr);
}

extern "C" int mobi_gemm_plan(const mobi_gemm_args* a, int32_t* tile_n, int32_t* pair, int32_t* persistent) {
    int32_t plan[3] = {0, 0, 0};
    if (gemm_impl(a, nullptr, plan)) return 1;
    if (tile_n) *tile_n = plan[0];
    if (pair) *pair = plan[1];
    if (persistent) *persistent = plan[2];
    return 0;
}
