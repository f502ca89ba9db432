// Persistent tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[M,N] = A[M,K] . B[N,K]^T     bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM walks output tiles (128 x BN, n-tile fastest so that the CTAs running at the same time share the
// same A rows through L2).  Warp roles (320 threads):
//   warp 0      TMA producer: 128x64 A tiles / BNx64 B tiles (128B swizzle) through a STAGES-deep ring that keeps
//               running across tile boundaries (the loads of tile i+1 start while tile i is still in the MMA);
//   warp 1      MMA issuer: tcgen05.mma M=128 N=BN K=16, accumulating into one of TWO TMEM accumulator stages, so
//               the epilogue of tile i overlaps the main loop of tile i+1; owns the TMEM allocation;
//   warps 2-9   epilogue: two warps per 32-lane TMEM group, each takes half of the tile's columns in 32-column
//               chunks: tcgen05.ld (thread = row) -> XOR-swizzled shared-memory transpose -> thread = 4 consecutive
//               columns, 8 lanes per row.  All global traffic of the epilogue (bias, per-image row bias, residual
//               in fp32/bf16, output in fp32/bf16/head-split layouts) is then 16-byte vectors, 128 contiguous bytes
//               per row, and the residual of a chunk is prefetched before its accumulator is read.
// With `conv` set the A tile of filter tap (kh,kw) is a shifted 4-D TMA box over the NHWC image (zero-filled halo).
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int G2_THREADS = 320;
constexpr int G2_STAGING_BYTES = 8 * 4096;  // 8 epilogue warps x (32 rows x 128 B)

// PM: 0 = one CTA per tile, 1 = CTA pair (256 x BN), 2 = two pairs per cluster with multicast A, 3 = WIDE pair: a
// 256 x 2BN tile per pair as two N = BN MMAs per k-step that share the A tile (see the kernel).
template <int BN, int PM = 0>
struct Gemm2Cfg {
    static constexpr bool WIDE = PM == 3;
    static constexpr int B_ROWS = WIDE ? BN : (PM >= 1 ? BN / 2 : BN);  // rows of B this CTA stages per k-block
    static constexpr int B_TILE_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    // WIDE: three BN-column slots that the tiles take two at a time, round robin (BN = 160: columns 0 / 160 / 320; a
    // 160-column accumulator may start at any multiple of 32 columns: measured, and what the wide-tile tests exercise)
    static constexpr int ACC_STRIDE = WIDE ? BN : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    static constexpr int TMEM_COLS = WIDE ? 512 : 2 * ACC_STRIDE;
    static constexpr int BUDGET = 227 * 1024 - 1024 - G2_STAGING_BYTES - 512;
    static constexpr int STAGES = (BUDGET / STAGE_BYTES) > 8 ? 8 : (BUDGET / STAGE_BYTES);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + G2_STAGING_BYTES + 512 + 1024;
    static constexpr int CHUNKS = BN / 32;  // 32-column epilogue chunks (BN is a multiple of 32)
};

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below bf16 output resolution): 2 MUFU + ~10 FMA
__device__ __forceinline__ float erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float y = 1.0f - poly * t * __expf(-ax * ax);
    return copysignf(y, x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752f)); }
// GELU(gate) * value for two (value, gate) pairs at once on the packed-fp32 pipe:
//   gelu(x) = x * Phi(x),  Phi(x) ~= 0.5 * (1 + tanh(x * (k + a x^2 + b x^4)))
// with (k, a, b) = (0.797507884, 0.0370056460, -0.000351516789) fitted to the exact erf form: max |error| of gelu over
// the real line 2.5e-5 (the textbook tanh-GELU is 4.7e-4 off); x^2 is clamped at 64, where tanh has long saturated.
// tanh.approx adds <= 2^-11 relative, i.e. everything stays ~8x below the bf16 resolution of the output.
// 1 MUFU + 5.5 issue slots per output instead of 2 MUFU + ~18 for the erf formula: with K = 320 the GEGLU epilogue
// would otherwise cost more cycles than the tile's MMAs.
__device__ __forceinline__ uint32_t geglu_pair_bf16(float v0, float g0, float v1, float g1) {
    const uint64_t x = pack2(g0, g1);
    float q0, q1;
    unpack2(mul2(x, x), q0, q1);
    const uint64_t x2 = pack2(fminf(q0, 64.f), fminf(q1, 64.f));
    uint64_t pl = fma2(pack2(-0.000351516789f, -0.000351516789f), x2, pack2(0.0370056460f, 0.0370056460f));
    pl = fma2(pl, x2, pack2(0.797507884f, 0.797507884f));
    float u0, u1;
    unpack2(mul2(pl, x), u0, u1);
    const uint64_t th = pack2(tanh_approx(u0), tanh_approx(u1));
    const uint64_t phi = fma2(th, pack2(0.5f, 0.5f), pack2(0.5f, 0.5f));
    float o0, o1;
    unpack2(mul2(mul2(x, phi), pack2(v0, v1)), o0, o1);
    return pack_bf16x2(o0, o1);
}
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

struct EpiRow {          // per-lane view of the 8 rows this lane serves in the transposed domain
    long long off[8];    // element offset of the row in the output (layout dependent)
    int rb[8];           // row-bias row offset (elements)
    unsigned ok;         // bit i: row i exists (m < M)
};

// The PLAIN epilogue of the hot f32 paths (attention / feed-forward output projections with their residual, proj_in /
// proj_out, every convolution): whole tiles (M % 128 == 0, N % 32 == 0), f32 output, optional f32 residual, bias, one
// row-bias row per 32-row group, optional column statistics; no activation, no atomics, no row segments.  Everything the
// general version decides per row is decided once per tile here, so the 8-row loop is straight-line code: the general
// version spent a third of its samples on branch resolution, reconvergence and instruction fetch
// (profiles/r02/ncu_gemm_l2_bound.md).
template <int BN>
__device__ __forceinline__ void gemm2_epilogue_plain_fast(const GemmParams& p, uint32_t acc_tmem, float* stg, int m_tile,
                                                          int n_tile, int lg, int half, int lane, long long batch_off,
                                                          int flip) {
    using Cfg = Gemm2Cfg<BN>;
    if (m_tile * BM >= p.M) return;   // the second CTA of a pair on an odd number of 128-row tiles
    const int u = lane & 7, rsub = lane >> 3;
    const int m_base = m_tile * BM + lg * 32;
    // row segments (out_seg % 128 == 0 here): the 128 rows of a tile stay consecutive in the output
    long long mo = m_base + rsub;
    if (p.out_seg > 0) {
        const int sg = (m_tile * BM) / (int)p.out_seg;
        mo += (long long)sg * (p.out_seg_stride - p.out_seg) + p.out_seg_offset;
    }
    const long long row0 = mo * p.ldo + batch_off + 4 * u;   // row `it` of this lane: + it * step
    const long long step = 4ll * p.ldo;
    float* outp = reinterpret_cast<float*>(p.out) + row0;
    const float* resp = p.residual ? reinterpret_cast<const float*>(p.residual) + row0 : nullptr;
    const float* rbp = p.row_bias ? p.row_bias + (long long)(m_base / p.rows_per_group) * p.ld_row_bias + 4 * u : nullptr;
    const float* biasp = p.bias ? p.bias + 4 * u : nullptr;
    const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
    uint8_t* srow = reinterpret_cast<uint8_t*>(stg) + lane * 128;
    const int split = (Cfg::CHUNKS + 1 - flip) / 2;
    const int c_begin = half == 0 ? 0 : split;
    const int c_end = half == 0 ? split : Cfg::CHUNKS;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
        const int n0 = n_tile * BN + c * 32;
        if (n0 >= p.N) break;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (biasp) b4 = __ldg(reinterpret_cast<const float4*>(biasp + n0));
        if (rbp) {
            const float4 rb = __ldg(reinterpret_cast<const float4*>(rbp + n0));
            b4.x += rb.x; b4.y += rb.y; b4.z += rb.z; b4.w += rb.w;
        }
        float4 res[8];
        if (resp) {
#pragma unroll
            for (int it = 0; it < 8; ++it) res[it] = *reinterpret_cast<const float4*>(resp + n0 + it * step);
        } else {
#pragma unroll
            for (int it = 0; it < 8; ++it) res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t r[32];
        tmem_ld32(acc_tmem + lane_addr + c * 32, r);
        tmem_ld_wait32(r);
#pragma unroll
        for (int unit = 0; unit < 8; ++unit)
            *reinterpret_cast<uint4*>(srow + ((unit ^ (lane & 7)) << 4)) =
                make_uint4(r[4 * unit], r[4 * unit + 1], r[4 * unit + 2], r[4 * unit + 3]);
        __syncwarp();
        float4 v[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            v[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stg) + rr * 128 + ((u ^ (rr & 7)) << 4));
        }
        uint64_t cs01 = 0, cs23 = 0, cq01 = 0, cq23 = 0;
        const uint64_t b01 = pack2(b4.x, b4.y), b23 = pack2(b4.z, b4.w);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const uint64_t w01 = add2(add2(pack2(v[it].x, v[it].y), b01), pack2(res[it].x, res[it].y));
            const uint64_t w23 = add2(add2(pack2(v[it].z, v[it].w), b23), pack2(res[it].z, res[it].w));
            cs01 = add2(cs01, w01);
            cs23 = add2(cs23, w23);
            cq01 = fma2(w01, w01, cq01);
            cq23 = fma2(w23, w23, cq23);
            float4 w;
            unpack2(w01, w.x, w.y);
            unpack2(w23, w.z, w.w);
            *reinterpret_cast<float4*>(outp + n0 + it * step) = w;
        }
        if (p.colstats) {
            float cst[8];
            unpack2(cs01, cst[0], cst[1]);
            unpack2(cs23, cst[2], cst[3]);
            unpack2(cq01, cst[4], cst[5]);
            unpack2(cq23, cst[6], cst[7]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cst[i] += __shfl_xor_sync(0xffffffffu, cst[i], 8);
                cst[i] += __shfl_xor_sync(0xffffffffu, cst[i], 16);
            }
            if (rsub == 0) {
                const long long groups32 = (p.M + 31) / 32;
                float* dst = p.colstats + (long long)(m_base >> 5) * p.N + n0 + 4 * u;
                *reinterpret_cast<float4*>(dst) = make_float4(cst[0], cst[1], cst[2], cst[3]);
                *reinterpret_cast<float4*>(dst + groups32 * p.N) = make_float4(cst[4], cst[5], cst[6], cst[7]);
            }
        }
        __syncwarp();
    }
}

// MC = epilogue class fixed at compile time (the runtime `p.mode` branches, the integer divisions of the head-split
// layouts and their registers disappear from the other classes): 0 = any mode (runtime), 1 = PLAIN, 2 = GEGLU2,
// 3 = head-split layouts without a transposed part (HEADS, QKV_ROW, KV_ROW).
template <int BN, int MC>
__device__ __forceinline__ void gemm2_epilogue(const GemmParams& p, uint32_t acc_tmem, float* stg, int m_tile,
                                               int n_tile, int lg, int half, int lane, long long batch_off, int flip) {
    using Cfg = Gemm2Cfg<BN>;
    if constexpr (MC == 1) {
        const bool fast = p.out_f32 && !p.atomic_out && p.act == 0 && (p.out_seg % BM) == 0 && (p.M % BM) == 0 && (p.N % 32) == 0 &&
                          (p.out_seg == 0 || p.colstats == nullptr) &&
                          (p.residual == nullptr || p.res_f32) && (p.row_bias == nullptr || (p.rows_per_group % 32) == 0);
        if (fast) {
            gemm2_epilogue_plain_fast<BN>(p, acc_tmem, stg, m_tile, n_tile, lg, half, lane, batch_off, flip);
            return;
        }
    }
    const int mode = MC == 1 ? MOBI_EPI_PLAIN : (MC == 2 ? MOBI_EPI_GEGLU2 : p.mode);
    const int u = lane & 7;        // 4-column unit inside a 32-column chunk
    const int rsub = lane >> 3;    // row inside a group of 4
    const int m_base = m_tile * BM + lg * 32;
    const int inner = p.heads * p.head_dim;
    const bool head_mode = MC == 3 || (mode >= MOBI_EPI_HEADS && mode <= MOBI_EPI_KV) || mode >= MOBI_EPI_QKV_ROW;
    const int nparts_last =
        MC == 3 ? -1 : (mode == MOBI_EPI_QKV ? 2 : (mode == MOBI_EPI_KV ? 1 : (mode == MOBI_EPI_HEADS_T ? 0 : -1)));
    // exact integer division by a small runtime divisor without the ~25-instruction IDIV sequence:
    // floor((x + 0.5) / d) in fp32 is exact while x < 2^21 (the fraction stays >= 0.5/d away from an integer)
    const float inv_inner = head_mode ? 1.0f / (float)inner : 0.f;
    const float inv_hd = head_mode ? 1.0f / (float)p.head_dim : 0.f;
    const bool tile_in_one_row = head_mode && (p.tokens % BM) == 0;  // all 128 rows of the tile share the batch row
    const int b_tile = tile_in_one_row ? (m_tile * BM) / p.tokens : 0;

    EpiRow er;
    er.ok = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int m = m_base + it * 4 + rsub;
        if (m < p.M) er.ok |= 1u << it;
        if (head_mode) {
            const int b = tile_in_one_row ? b_tile : m / p.tokens;
            const int t = m - b * p.tokens;
            er.off[it] = ((long long)b * p.heads * p.tokens + t) * p.head_dim;  // + (h*tokens*d + dd) per column
        } else {
            long long mo = m;
            if (p.out_seg > 0) {
                const int sg = m / (int)p.out_seg;
                mo = (long long)sg * p.out_seg_stride + p.out_seg_offset + (m - sg * (int)p.out_seg);
            }
            er.off[it] = mo * p.ldo + batch_off;
        }
        er.rb[it] = p.row_bias ? (m / p.rows_per_group) * (int)p.ld_row_bias : 0;
    }
    // row-domain identity (thread = TMEM lane = row), needed for the transposed-V layout
    const int m_own = m_base + lane;
    long long vt_row = 0;
    if (nparts_last >= 0 && m_own < p.M) {
        const int b = m_own / p.tokens, t = m_own - b * p.tokens;
        vt_row = (long long)b * inner * p.tokens + t;
    }
    const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
    uint8_t* srow = reinterpret_cast<uint8_t*>(stg) + lane * 128;  // row-domain staging row of this thread

    // With an odd chunk count (BN = 160: 5) the extra chunk alternates between the two column halves from tile to tile
    // (`flip`), so both warps of a lane group do 5 chunks per two tiles instead of one doing 6 and the other waiting.
    const int split = (Cfg::CHUNKS + 1 - flip) / 2;
    const int c_begin = half == 0 ? 0 : split;
    const int c_end = half == 0 ? split : Cfg::CHUNKS;
    // The accumulator chunk of iteration c + 1 is requested as soon as the registers of chunk c have been staged, so the
    // TMEM read runs under the transposed-domain work (shared-memory reads, activation, global stores) of chunk c.
    // Only for the bf16-output classes (GEGLU2, head-split): the fp32 / residual class needs the 32 registers for its
    // residual prefetch and spills with both (measured 3-8 % slower).
    constexpr bool TMEM_AHEAD = MC >= 2;
    uint32_t r[32];
    if (TMEM_AHEAD && c_begin < c_end && n_tile * BN + c_begin * 32 < p.N)
        tmem_ld32(acc_tmem + lane_addr + c_begin * 32, r);
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
        const int n0 = n_tile * BN + c * 32;
        if (n0 >= p.N) break;
        const bool more = c + 1 < c_end && n0 + 32 < p.N;
        const int n = n0 + 4 * u;             // first of this lane's 4 columns (transposed domain)
        const bool col_ok = n < p.N;          // N % 4 == 0: the unit is entirely inside or outside
        // bias of this lane's 4 columns: issued before the accumulator load so its latency hides under it
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mode != MOBI_EPI_GEGLU && p.bias && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        // ---- prefetch the residual of this chunk (transposed domain)
        float4 res[8];
        if (p.residual != nullptr && mode == MOBI_EPI_PLAIN) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col_ok && ((er.ok >> it) & 1)) {
                    if (p.res_f32) {
                        res[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + er.off[it] + n);
                    } else {
                        const uint2 rw = *reinterpret_cast<const uint2*>(
                            reinterpret_cast<const __nv_bfloat16*>(p.residual) + er.off[it] + n);
                        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw.x));
                        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw.y));
                        res[it] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        // ---- accumulator chunk -> registers (thread = row): requested one iteration ago
        if (!TMEM_AHEAD) tmem_ld32(acc_tmem + lane_addr + c * 32, r);
        tmem_ld_wait32(r);
        int out_units = 8;  // 16-byte units per staged row
        if (mode == MOBI_EPI_GEGLU) {
            // [8 value | 8 gate] groups: activation in the row domain, 16 outputs per chunk
            out_units = 4;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float val = __uint_as_float(r[16 * g + j]);
                    float gate = __uint_as_float(r[16 * g + 8 + j]);
                    if (p.bias) {
                        val += __ldg(p.bias + n0 + 16 * g + j);
                        gate += __ldg(p.bias + n0 + 16 * g + 8 + j);
                    }
                    o[j] = val * gelu_fast(gate);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int unit = 2 * g + q;
                    *reinterpret_cast<float4*>(srow + ((unit ^ (lane & 7)) << 4)) =
                        make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
                }
            }
        } else {
            if (nparts_last >= 0 && m_own < p.M) {
                // transposed-V columns go straight from the row domain: 32 consecutive tokens per store
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int ng = n0 + 8 * g;
                    if (ng < p.N && ng / inner == nparts_last) {
                        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(
                            nparts_last == 2 ? p.out3 : (nparts_last == 1 ? p.out2 : p.out));
                        __nv_bfloat16* o = base + vt_row + (long long)(ng - nparts_last * inner) * p.tokens;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float val = __uint_as_float(r[8 * g + j]);
                            if (p.bias) val += __ldg(p.bias + ng + j);
                            o[(long long)j * p.tokens] = __float2bfloat16(val);
                        }
                    }
                }
            }
#pragma unroll
            for (int unit = 0; unit < 8; ++unit)
                *reinterpret_cast<uint4*>(srow + ((unit ^ (lane & 7)) << 4)) =
                    make_uint4(r[4 * unit], r[4 * unit + 1], r[4 * unit + 2], r[4 * unit + 3]);
        }
        __syncwarp();
        if (TMEM_AHEAD && more) tmem_ld32(acc_tmem + lane_addr + (c + 1) * 32, r);
        // ---- transposed domain
        if (mode == MOBI_EPI_GEGLU2) {
            // (value, gate) column pairs: two outputs per lane, no cross-lane traffic.  Branch-free math over the 8
            // rows of this lane (16 independent GELUs in flight), only the stores are predicated.
            uint32_t o[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rsub;
                const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stg) + rr * 128 +
                                                                  ((u ^ (rr & 7)) << 4));
                o[it] = geglu_pair_bf16(v.x + b4.x, v.y + b4.y, v.z + b4.z, v.w + b4.w);
            }
#pragma unroll
            for (int it = 0; it < 8; ++it)
                if (col_ok && ((er.ok >> it) & 1))
                    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out) + er.off[it] + (n >> 1)) = o[it];
        } else if (out_units == 8) {
            // head-split column part
            long long coloff = 0;
            __nv_bfloat16* hbase = nullptr;
            bool is_vt = false;
            if (head_mode && col_ok) {
                const int which = (int)(((float)n + 0.5f) * inv_inner);
                const int nn = n - which * inner;
                is_vt = (which == nparts_last);
                const int h = (int)(((float)nn + 0.5f) * inv_hd), dd = nn - h * p.head_dim;
                coloff = (long long)h * p.tokens * p.head_dim + dd;
                hbase = reinterpret_cast<__nv_bfloat16*>(which == 0 ? p.out : (which == 1 ? p.out2 : p.out3));
            }
            // all 8 rows of this lane leave the staging buffer before any of them is used (one shared-memory latency per
            // chunk instead of one per row; the per-row predicates only guard the global accesses)
            float4 v[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rsub;
                v[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stg) + rr * 128 +
                                                         ((u ^ (rr & 7)) << 4));
            }
            // column statistics of the final values for the GroupNorm that consumes this output (mobi_gemm_args.colstats)
            uint64_t cs01 = 0, cs23 = 0, cq01 = 0, cq23 = 0;  // packed (0.f, 0.f)
            if (head_mode) {
                if (col_ok && !is_vt) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        uint2 pk;
                        pk.x = pack_bf16x2(v[it].x + b4.x, v[it].y + b4.y);
                        pk.y = pack_bf16x2(v[it].z + b4.z, v[it].w + b4.w);
                        if ((er.ok >> it) & 1) *reinterpret_cast<uint2*>(hbase + er.off[it] + coloff) = pk;
                    }
                }
            } else {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const bool live = col_ok && ((er.ok >> it) & 1);
                    float4 w = v[it];
                    w.x += b4.x; w.y += b4.y; w.z += b4.z; w.w += b4.w;
                    if (p.row_bias && live) {
                        const float4 rb = __ldg(reinterpret_cast<const float4*>(p.row_bias + er.rb[it] + n));
                        w.x += rb.x; w.y += rb.y; w.z += rb.z; w.w += rb.w;
                    }
                    if (p.act == 1) {
                        w.x = silu_fast(w.x); w.y = silu_fast(w.y); w.z = silu_fast(w.z); w.w = silu_fast(w.w);
                    } else if (p.act == 2) {
                        w.x = gelu_fast(w.x); w.y = gelu_fast(w.y); w.z = gelu_fast(w.z); w.w = gelu_fast(w.w);
                    }
                    if (p.residual) {
                        w.x += res[it].x; w.y += res[it].y; w.z += res[it].z; w.w += res[it].w;
                    }
                    if (!live) continue;
                    if (MC <= 1 && p.colstats) {
                        const uint64_t v01 = pack2(w.x, w.y), v23 = pack2(w.z, w.w);
                        cs01 = add2(cs01, v01);
                        cs23 = add2(cs23, v23);
                        cq01 = fma2(v01, v01, cq01);
                        cq23 = fma2(v23, v23, cq23);
                    }
                    if (p.out_f32) {
                        float* dst = reinterpret_cast<float*>(p.out) + er.off[it] + n;
                        if (p.atomic_out) {
                            atomicAdd(dst, w.x);
                            atomicAdd(dst + 1, w.y);
                            atomicAdd(dst + 2, w.z);
                            atomicAdd(dst + 3, w.w);
                        } else {
                            *reinterpret_cast<float4*>(dst) = w;
                        }
                    } else {
                        uint2 pk;
                        pk.x = pack_bf16x2(w.x, w.y);
                        pk.y = pack_bf16x2(w.z, w.w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + er.off[it] + n) = pk;
                    }
                }
            }
            if (MC <= 1 && p.colstats) {
                // the 4 lanes that share this 4-column unit (rsub = 0..3) hold the 32 rows of the warp's lane group
                float c[8];
                unpack2(cs01, c[0], c[1]);
                unpack2(cs23, c[2], c[3]);
                unpack2(cq01, c[4], c[5]);
                unpack2(cq23, c[6], c[7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    c[i] += __shfl_xor_sync(0xffffffffu, c[i], 8);
                    c[i] += __shfl_xor_sync(0xffffffffu, c[i], 16);
                }
                if (rsub == 0 && col_ok && m_base < p.M) {
                    const long long groups32 = (p.M + 31) / 32;
                    float* dst = p.colstats + (long long)(m_base >> 5) * p.N + n;
                    *reinterpret_cast<float4*>(dst) = make_float4(c[0], c[1], c[2], c[3]);
                    *reinterpret_cast<float4*>(dst + groups32 * p.N) = make_float4(c[4], c[5], c[6], c[7]);
                }
            }
        } else {
            // GEGLU: 16 bf16 outputs per row: lane -> (row = it*8 + lane/4, unit = lane%4)
            const int u4 = lane & 3, r8 = lane >> 2;
            const int no = n0 / 2 + 4 * u4;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + r8;
                const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stg) + rr * 128 +
                                                                  ((u4 ^ (rr & 7)) << 4));
                const int m = m_base + rr;
                if (m < p.M && no < p.N / 2) {
                    uint2 pk;
                    pk.x = pack_bf16x2(v.x, v.y);
                    pk.y = pack_bf16x2(v.z, v.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)m * p.ldo + no) = pk;
                }
            }
        }
        __syncwarp();
    }
}

// PAIR: the kernel runs as 2-CTA clusters (cta_group::2).  A pair owns a 256 x BN tile: CTA r computes rows
// [m0 + 128 r, +128) into its own TMEM from its own A tile and HALF of the B tile (BN / 2 rows) — per MMA an SM reads
// 4 KB of A + BN/2 * 32 B of B from shared memory instead of 4 KB + BN * 32 B, which lifts the 128 x 160 tiles of the
// convolutions off the shared-memory read port.  The leader's warp 1 issues every MMA; TMA completions of both CTAs are
// counted on the leader's `full` barriers, tcgen05.commit multicasts `empty` / `tfull` to both CTAs, and the epilogue
// warps of both CTAs arrive on the leader's `tempty`.
// Batched operand loads: one batch index z, or (z % inner, z / inner) over 4-D maps when the batch has two levels
// (heads inside batch rows: head columns of token-major matrices have no single batch stride).
__device__ __forceinline__ void tma_load_batched(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int z, int inner) {
    if (inner > 0) tma_load_4d(dst, tm, bar, x, y, z % inner, z / inner);
    else tma_load_3d(dst, tm, bar, x, y, z);
}
__device__ __forceinline__ void tma_load_batched_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int z, int inner) {
    if (inner > 0) tma_load_4d_pair(dst, tm, bar, x, y, z % inner, z / inner);
    else tma_load_3d_pair(dst, tm, bar, x, y, z);
}

//
// QUAD (PM = 2): clusters of FOUR CTAs = two pairs on the same 256 rows and on neighbouring n-tiles (2j, 2j + 1).  Both pairs
// need the same A tiles, so each CTA loads only HALF of its 128 x 64 A tile (64 rows, 8 KB) and multicasts it to its
// counterpart in the other pair (rank ^ 2): per k-block a CTA pulls 8 KB of A + BN / 2 rows of B through L2 instead of
// 16 KB + BN / 2 rows.  These kernels sit at ~80 % of the L2 -> SM throughput cap with the tensor pipe half idle
// (profiles/r02/ncu_gemm_l2_bound.md); A is 60 % of that traffic at BN = 160.  The ring of the two pairs moves in
// lock-step: every `empty` barrier counts the commits of BOTH leaders (a stage is refilled by two CTAs), everything else
// (full / tfull / tempty, TMEM) stays per pair.
template <int BN, int MC, int PM>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmR, const GemmParams p, const int m_tiles, const int n_tiles) {
    constexpr bool PAIR = PM >= 1;
    constexpr bool QUAD = PM == 2;
    constexpr bool WIDE = PM == 3;
    using Cfg = Gemm2Cfg<BN, PM>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // align by OFFSET so the pointer keeps its shared address space (LDS/STS instead of generic LD/ST in the epilogue)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_TILE_BYTES;
    float* staging = reinterpret_cast<float*>(sB + STAGES * Cfg::B_TILE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + G2_STAGING_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;   // 2 (3 accumulator slots for WIDE)
    uint64_t* tempty_bar = tfull_bar + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 3);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int cl_rank = PAIR ? (int)cluster_ctarank() : 0;   // rank in the cluster (0..1, or 0..3 for a quad)
    const int cta_rank = cl_rank & 1;                          // rank in the pair: 0 = leader
    const int quad_pair = QUAD ? cl_rank >> 1 : 0;             // which pair of the quad = which of its two n-tiles
    const int worker = QUAD ? (int)(blockIdx.x >> 2) : (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);  // CTA / pair / quad
    const int n_workers = QUAD ? (int)(gridDim.x >> 2) : (PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x);
    const int n_cols = QUAD ? n_tiles >> 1 : n_tiles;          // tile columns a worker walks (a quad takes two n-tiles)

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], QUAD ? 2 : 1);
            }
            for (int s = 0; s < 3; ++s) {
                mbar_init(&tfull_bar[s], 1);
                // one elected arrive per epilogue warp (of both CTAs)
                mbar_init(&tempty_bar[s], PAIR ? 16 : 8);
            }
            fence_barrier_init();
        }
    } else if (warp == 1) {
        if (PAIR) {
            tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();  // barriers of BOTH CTAs initialised before any remote arrive / TMA signal
    else __syncthreads();
    tc_fence_after();
    pdl_wait();     // everything above overlapped the tail of the previous kernel; its results are needed from here on
    pdl_trigger();
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = p.num_k_blocks;
    // tiles are 128 x BN (or 256 x BN per pair: m_tiles then counts 256-row tiles)
    const int per_batch = m_tiles * n_cols;
    const int total = per_batch * (p.batch > 1 ? p.batch : 1);

    if (warp == 0) {
        if (elect_one()) {
            // ---------------- TMA producer
            uint32_t it = 0;
            for (int tile = worker; tile < total; tile += n_workers) {
                const int z = tile / per_batch, rem = tile - z * per_batch;
                const int m_tile = PAIR ? (rem / n_cols) * 2 + cta_rank : rem / n_cols;   // this CTA's 128-row tile
                const int n_tile = QUAD ? (rem % n_cols) * 2 + quad_pair : rem % n_cols;
                // this CTA's rows of B (WIDE: its BN / 2 rows of the tile's first N-half; the second half is BN further)
                const int b_row0 = n_tile * (WIDE ? 2 * BN : BN) + (PAIR ? cta_rank * (BN / 2) : 0);
                const int a_half = QUAD ? quad_pair * (BM / 2) : 0;   // quad: the 64 rows of the A tile this CTA fetches
                const uint16_t a_mask = (uint16_t)((1u << cl_rank) | (1u << (cl_rank ^ 2)));
                // The residual of this CTA's 128 x BN output box starts its way from HBM to L2 now, one to two tile periods
                // before the epilogue warps read it: their loads then wait for an L2 hit instead of a DRAM access (the
                // narrow f32 GEMMs were bound by exactly that latency, profiles/r02/ncu_gemm_l2_bound.md).
                if (p.res_prefetch) {
                    long long mo = (long long)m_tile * BM;
                    if (p.out_seg > 0) mo += (mo / p.out_seg) * (p.out_seg_stride - p.out_seg) + p.out_seg_offset;
                    tma_prefetch_l2_2d(&tmR, n_tile * (WIDE ? 2 * BN : BN), (int)mo);
                    if (WIDE && n_tile * 2 * BN + BN < p.N) tma_prefetch_l2_2d(&tmR, n_tile * 2 * BN + BN, (int)mo);
                }
                int x0 = 0, y0 = 0, n0 = 0;
                if (p.conv) {
                    const long long pix = (long long)m_tile * BM + a_half;
                    x0 = (int)(pix % p.W);
                    y0 = (int)((pix / p.W) % p.H);
                    n0 = (int)(pix / ((long long)p.W * p.H));
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* dA = sA + s * A_TILE_BYTES;
                    uint8_t* dB = sB + s * Cfg::B_TILE_BYTES;
                    if (PAIR) {
                        // both CTAs' bytes are counted on the leader's barrier; only the leader arms it
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_TILE_BYTES + Cfg::B_TILE_BYTES));
                        if (WIDE) {
                            // the A tile once, the CTA's B rows of both N-halves (two boxes of BN / 2 rows)
                            constexpr int HB = (BN / 2) * BK * 2;
                            if (p.conv) {
                                const int tap = kb / p.cblocks;
                                const int cb = kb - tap * p.cblocks;
                                const int kh = tap / p.KW;
                                const int kw = tap - kh * p.KW;
                                tma_load_4d_pair(dA, &tmA, &full_bar[s], cb * BK, x0 + kw - p.pad_w, y0 + kh - p.pad_h, n0);
                                tma_load_2d_pair(dB, &tmB, &full_bar[s], tap * p.C + cb * BK, b_row0);
                                tma_load_2d_pair(dB + HB, &tmB, &full_bar[s], tap * p.C + cb * BK, b_row0 + BN);
                            } else {
                                tma_load_2d_pair(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM);
                                tma_load_2d_pair(dB, &tmB, &full_bar[s], kb * BK, b_row0);
                                tma_load_2d_pair(dB + HB, &tmB, &full_bar[s], kb * BK, b_row0 + BN);
                            }
                        } else if (QUAD) {
                            // half of the A tile (the A map's box is 64 rows), written into this CTA and its counterpart
                            uint8_t* dAh = dA + a_half * 128;
                            if (p.conv) {
                                const int tap = kb / p.cblocks;
                                const int cb = kb - tap * p.cblocks;
                                const int kh = tap / p.KW;
                                const int kw = tap - kh * p.KW;
                                tma_load_4d_pair_mc(dAh, &tmA, &full_bar[s], a_mask, cb * BK, x0 + kw - p.pad_w,
                                                    y0 + kh - p.pad_h, n0);
                                tma_load_2d_pair(dB, &tmB, &full_bar[s], tap * p.C + cb * BK, b_row0);
                            } else {
                                tma_load_2d_pair_mc(dAh, &tmA, &full_bar[s], a_mask, kb * BK, m_tile * BM + a_half);
                                tma_load_2d_pair(dB, &tmB, &full_bar[s], kb * BK, b_row0);
                            }
                        } else if (p.conv) {
                            const int tap = kb / p.cblocks;
                            const int cb = kb - tap * p.cblocks;
                            const int kh = tap / p.KW;
                            const int kw = tap - kh * p.KW;
                            tma_load_4d_pair(dA, &tmA, &full_bar[s], cb * BK, x0 + kw - p.pad_w, y0 + kh - p.pad_h, n0);
                            tma_load_2d_pair(dB, &tmB, &full_bar[s], tap * p.C + cb * BK, b_row0);
                        } else if (p.batch > 1) {
                            tma_load_batched_pair(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM, z, p.batch_inner);
                            tma_load_batched_pair(dB, &tmB, &full_bar[s], kb * BK, b_row0, z, p.batch_inner);
                        } else {
                            tma_load_2d_pair(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM);
                            tma_load_2d_pair(dB, &tmB, &full_bar[s], kb * BK, b_row0);
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_bar[s], A_TILE_BYTES + Cfg::B_TILE_BYTES);
                    if (p.a_mn | p.b_mn) {
                        // MN-major operands: the tile is [64 k-rows][64 MN elements = 128 B] boxes, one per 64 rows of
                        // the tile; K-major operands keep their single [rows][64 k] box
                        if (p.a_mn) {
                            for (int c = 0; c < BM / 64; ++c) {
                                if (p.batch > 1) tma_load_batched(dA + c * 8192, &tmA, &full_bar[s], m_tile * BM + c * 64, kb * BK, z, p.batch_inner);
                                else tma_load_2d(dA + c * 8192, &tmA, &full_bar[s], m_tile * BM + c * 64, kb * BK);
                            }
                        } else if (p.batch > 1) {
                            tma_load_batched(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM, z, p.batch_inner);
                        } else {
                            tma_load_2d(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM);
                        }
                        if (p.b_mn) {
                            for (int c = 0; c < BN / 64; ++c) {
                                if (p.batch > 1) tma_load_batched(dB + c * 8192, &tmB, &full_bar[s], b_row0 + c * 64, kb * BK, z, p.batch_inner);
                                else tma_load_2d(dB + c * 8192, &tmB, &full_bar[s], b_row0 + c * 64, kb * BK);
                            }
                        } else if (p.batch > 1) {
                            tma_load_batched(dB, &tmB, &full_bar[s], kb * BK, b_row0, z, p.batch_inner);
                        } else {
                            tma_load_2d(dB, &tmB, &full_bar[s], kb * BK, b_row0);
                        }
                    } else if (p.conv) {
                        const int tap = kb / p.cblocks;
                        const int cb = kb - tap * p.cblocks;
                        const int kh = tap / p.KW;
                        const int kw = tap - kh * p.KW;
                        tma_load_4d(dA, &tmA, &full_bar[s], cb * BK, x0 + kw - p.pad_w, y0 + kh - p.pad_h, n0);
                        tma_load_2d(dB, &tmB, &full_bar[s], tap * p.C + cb * BK, b_row0);
                    } else if (p.batch > 1) {
                        tma_load_batched(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM, z, p.batch_inner);
                        tma_load_batched(dB, &tmB, &full_bar[s], kb * BK, b_row0, z, p.batch_inner);
                    } else {
                        tma_load_2d(dA, &tmA, &full_bar[s], kb * BK, m_tile * BM);
                        tma_load_2d(dB, &tmB, &full_bar[s], kb * BK, b_row0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (cta_rank == 0 && elect_one()) {
            // ---------------- MMA issuer (the leader CTA's for a pair)
            const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BN) | (p.a_mn ? 1u << 15 : 0u) | (p.b_mn ? 1u << 16 : 0u);
            uint32_t it = 0, lt = 0;
            for (int tile = worker; tile < total; tile += n_workers, ++lt) {
                // WIDE: accumulator uses are numbered u = 2 lt + half and take slot u % 3; the (u / 3)-th use of a slot
                const uint32_t u0 = 2 * lt, u1 = 2 * lt + 1;
                const uint32_t as = WIDE ? u0 % 3 : (lt & 1);
                const uint32_t as1 = u1 % 3;
                mbar_wait(&tempty_bar[as], WIDE ? (((u0 / 3) & 1) ^ 1) : (((lt >> 1) & 1) ^ 1));  // the epilogue has drained it
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
                const uint32_t d_tmem1 = tmem_base + as1 * Cfg::ACC_STRIDE;
                if constexpr (WIDE) {
                    // The slot of the SECOND N-half is the one the previous tile's first half is still being drained
                    // from; the first half's slot has been free for a whole tile.  So the first-half MMAs start at once
                    // and run up to STAGES - 1 k-blocks ahead (their stages stay held) while the epilogue finishes; the
                    // second half catches up as soon as its slot is released.  Stages are committed after both halves.
                    constexpr uint64_t HB16 = ((BN / 2) * BK * 2) >> 4;
                    const uint32_t par1 = ((u1 / 3) & 1) ^ 1;
                    const uint32_t it0 = it;
                    bool ready1 = false;
                    int done1 = 0;
                    for (int kb = 0; kb < nkb; ++kb, ++it) {
                        const int s = it % STAGES;
                        mbar_wait(&full_bar[s], (it / STAGES) & 1);
                        tc_fence_after();
                        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA + s * A_TILE_BYTES));
                        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sB + s * Cfg::B_TILE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (!ready1) {
                            if (kb - done1 >= STAGES - 2 || kb == nkb - 1) {
                                mbar_wait(&tempty_bar[as1], par1);
                                ready1 = true;
                            } else {
                                ready1 = mbar_try_wait(&tempty_bar[as1], par1);
                            }
                            if (ready1) tc_fence_after();
                        }
                        if (ready1) {
                            for (; done1 <= kb; ++done1) {
                                const int s1 = (it0 + done1) % STAGES;
                                const uint64_t ad1 = make_kmajor_sw128_desc(smem_u32(sA + s1 * A_TILE_BYTES));
                                const uint64_t bd1 = make_kmajor_sw128_desc(smem_u32(sB + s1 * Cfg::B_TILE_BYTES)) + HB16;
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_bf16_ss_pair(d_tmem1, ad1 + 2 * k, bd1 + 2 * k, idesc, (done1 | k) != 0 ? 1u : 0u);
                                umma_commit_pair(&empty_bar[s1]);
                            }
                        }
                    }
                    umma_commit_pair(&tfull_bar[as]);
                    umma_commit_pair(&tfull_bar[as1]);
                } else {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA + s * A_TILE_BYTES));
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sB + s * Cfg::B_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (PAIR) {
                            umma_bf16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        } else {
                            // MN-major: 16 k-rows = 2048 B into every 64-element chunk, chunks 8192 B apart (LBO)
                            const uint64_t ad = p.a_mn ? make_mnmajor_sw128_desc(smem_u32(sA + s * A_TILE_BYTES) + k * 2048, 8192)
                                                       : adesc + 2 * k;
                            const uint64_t bd = p.b_mn ? make_mnmajor_sw128_desc(smem_u32(sB + s * Cfg::B_TILE_BYTES) + k * 2048, 8192)
                                                       : bdesc + 2 * k;
                            umma_bf16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if (QUAD) umma_commit_mask(&empty_bar[s], 0xF);   // the stage is refilled by CTAs of both pairs
                    else if (PAIR) umma_commit_pair(&empty_bar[s]);
                    else umma_commit(&empty_bar[s]);
                }
                if (QUAD) umma_commit_mask(&tfull_bar[as], (uint16_t)(3u << (2 * quad_pair)));
                else if (PAIR) umma_commit_pair(&tfull_bar[as]);
                else umma_commit(&tfull_bar[as]);
                (void)d_tmem1;
                }
            }
        }
    } else {
        // ---------------- epilogue warps 2..9: TMEM lane group = warp % 4, column half = (warp - 2) / 4
        const int lg = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stg = staging + (warp - 2) * 1024;
        uint32_t lt = 0;
        for (int tile = worker; tile < total; tile += n_workers, ++lt) {
            const int z = tile / per_batch, rem = tile - z * per_batch;
            const int m_tile = PAIR ? (rem / n_cols) * 2 + cta_rank : rem / n_cols;
            const int n_tile = QUAD ? (rem % n_cols) * 2 + quad_pair : rem % n_cols;
            const long long boff = p.batch_inner > 0 ? (long long)(z % p.batch_inner) * p.out_batch_stride +
                                                           (long long)(z / p.batch_inner) * p.out_batch2_stride
                                                     : (long long)z * p.out_batch_stride;
            if constexpr (WIDE) {
                // Both accumulators of the tile, the first N-half first and by ALL warps: the next tile's second half
                // reuses exactly that slot, so the MMA issuer gets it back after half an epilogue.  The odd chunk of
                // each half goes to the other warp of the lane group (3 + 2 chunks per warp and tile).
#pragma unroll 1
                for (int part = 0; part < 2; ++part) {
                    const uint32_t u = 2 * lt + part;                 // u-th accumulator use: slot u % 3, its (u / 3)-th use
                    const uint32_t as = u % 3;
                    mbar_wait(&tfull_bar[as], (u / 3) & 1);
                    tc_fence_after();
                    gemm2_epilogue<BN, MC>(p, tmem_base + as * Cfg::ACC_STRIDE, stg, m_tile, 2 * n_tile + part, lg, half, lane, boff,
                                           (int)((lt + part) & 1));
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
                }
            } else {
            const uint32_t as = lt & 1;
            mbar_wait(&tfull_bar[as], (lt >> 1) & 1);
            tc_fence_after();
            gemm2_epilogue<BN, MC>(p, tmem_base + as * Cfg::ACC_STRIDE, stg, m_tile, n_tile, lg, half, lane, boff,
                                   (Cfg::CHUNKS & 1) ? (int)(lt & 1) : 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(&tempty_bar[as]);
                else mbar_arrive(&tempty_bar[as]);
            }
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// Clusters of 4 that fit on the device at once (GPCs whose SM count is not a multiple of 4 leave SMs idle: 136 of 148
// SMs on this part), asked from the occupancy calculator once per kernel.
template <int BN, int MC>
static int launch_gemm2_quad(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p,
                             cudaStream_t stream) {
    using Cfg = Gemm2Cfg<BN, 1>;
    static int max_quads = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(G2_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_quads == 0) {
        MOBI_CUDA(cudaFuncSetAttribute(gemm2_kernel<BN, MC, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        cfg.gridDim = dim3(4 * (sm_count() / 4), 1, 1);
        int n = 0;
        MOBI_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<BN, MC, 2>, &cfg));
        MOBI_CHECK(n > 0, "mobi_gemm: no 4-CTA cluster of the persistent kernel fits on this device");
        max_quads = n;
    }
    const int m_tiles2 = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + BN - 1) / BN;
    MOBI_CHECK(n_tiles % 2 == 0, "mobi_gemm: 4-CTA clusters need an even number of n-tiles");
    const long long total = (long long)m_tiles2 * (n_tiles / 2) * (p.batch > 1 ? p.batch : 1);
    MOBI_CHECK(total < (1ll << 31), "mobi_gemm: too many tiles");
    const int quads = (int)(total < max_quads ? total : max_quads);
    cfg.gridDim = dim3(4 * quads, 1, 1);
    MOBI_CUDA(cudaLaunchKernelEx(&cfg, gemm2_kernel<BN, MC, 2>, tmA, tmB, tmR, p, m_tiles2, n_tiles));
    return 0;
}

// WIDE pairs: n_tiles counts 2 BN-wide tiles.
template <int BN, int MC>
static int launch_gemm2_wide(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p,
                             cudaStream_t stream) {
    using Cfg = Gemm2Cfg<BN, 3>;
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(gemm2_kernel<BN, MC, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int m_tiles2 = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + 2 * BN - 1) / (2 * BN);
    const long long total = (long long)m_tiles2 * n_tiles * (p.batch > 1 ? p.batch : 1);
    MOBI_CHECK(total < (1ll << 31), "mobi_gemm: too many tiles");
    const int pairs = (int)(total < sm_count() / 2 ? total : sm_count() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(G2_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    MOBI_CUDA(cudaLaunchKernelEx(&cfg, gemm2_kernel<BN, MC, 3>, tmA, tmB, tmR, p, m_tiles2, n_tiles));
    return 0;
}

template <int BN, int MC>
static int launch_gemm2_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p,
                             cudaStream_t stream) {
    if (p.pair == 3) {
        if constexpr (BN == 160) {
            return launch_gemm2_wide<BN, MC>(tmA, tmB, tmR, p, stream);
        } else {
            MOBI_CHECK(false, "mobi_gemm: wide pairs are built for tile_n = 160 (320-column tiles)");
        }
    }
    if (p.pair == 2) {
        if constexpr (MC == 1 && BN != 64) {
            return launch_gemm2_quad<BN, MC>(tmA, tmB, tmR, p, stream);
        } else {
            MOBI_CHECK(false, "mobi_gemm: 4-CTA clusters are built for the PLAIN epilogue with tile_n >= 128");
        }
    }
    using Cfg = Gemm2Cfg<BN, 1>;
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(gemm2_kernel<BN, MC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
        configured = true;
    }
    const int m_tiles2 = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + BN - 1) / BN;
    const long long total = (long long)m_tiles2 * n_tiles * (p.batch > 1 ? p.batch : 1);
    MOBI_CHECK(total < (1ll << 31), "mobi_gemm: too many tiles");
    const int pairs = (int)(total < sm_count() / 2 ? total : sm_count() / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(G2_THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    MOBI_CUDA(cudaLaunchKernelEx(&cfg, gemm2_kernel<BN, MC, 1>, tmA, tmB, tmR, p, m_tiles2, n_tiles));
    return 0;
}

template <int BN, int MC>
static int launch_gemm2_tm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p,
                           cudaStream_t stream) {
    if (p.pair) return launch_gemm2_pair<BN, MC>(tmA, tmB, tmR, p, stream);
    using Cfg = Gemm2Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(gemm2_kernel<BN, MC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
        configured = true;
    }
    const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
    const long long total = (long long)m_tiles * n_tiles * (p.batch > 1 ? p.batch : 1);
    MOBI_CHECK(total < (1ll << 31), "mobi_gemm: too many tiles");
    const int grid = (int)(total < sm_count() ? total : sm_count());
    MOBI_CUDA(launch_pdl(gemm2_kernel<BN, MC, 0>, dim3(grid), dim3(G2_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmR,
                         p, m_tiles, n_tiles));
    return 0;
}

template <int BN>
static int launch_gemm2_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p,
                          cudaStream_t stream) {
    if (p.mode == MOBI_EPI_PLAIN) return launch_gemm2_tm<BN, 1>(tmA, tmB, tmR, p, stream);
    if (p.mode == MOBI_EPI_GEGLU2) return launch_gemm2_tm<BN, 2>(tmA, tmB, tmR, p, stream);
    if (p.mode == MOBI_EPI_HEADS || p.mode == MOBI_EPI_QKV_ROW || p.mode == MOBI_EPI_KV_ROW)
        return launch_gemm2_tm<BN, 3>(tmA, tmB, tmR, p, stream);
    return launch_gemm2_tm<BN, 0>(tmA, tmB, tmR, p, stream);
}

bool gemm2_supported(const GemmParams& p) {
    if (p.N % 4 != 0) return false;
    if (p.mode == MOBI_EPI_PLAIN) {
        if (p.ldo % 4 != 0) return false;
        const uintptr_t align = p.out_f32 ? 15 : 7;
        if (reinterpret_cast<uintptr_t>(p.out) & align) return false;
        if (p.residual) {
            if (reinterpret_cast<uintptr_t>(p.residual) & (p.res_f32 ? 15 : 7)) return false;
        }
        if (p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15)) return false;
        if (p.row_bias && ((reinterpret_cast<uintptr_t>(p.row_bias) & 15) || p.ld_row_bias % 4 != 0)) return false;
        if (p.out_seg >= (1ll << 31)) return false;
    } else if (p.mode == MOBI_EPI_GEGLU) {
        if (p.N % 32 != 0 || p.ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(p.out) & 7)) return false;
    } else if (p.mode == MOBI_EPI_GEGLU2) {
        if (p.ldo % 2 != 0 || (reinterpret_cast<uintptr_t>(p.out) & 3)) return false;
        if (p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15)) return false;
    } else {
        if (p.head_dim % 8 != 0) return false;
        if ((reinterpret_cast<uintptr_t>(p.out) & 7) || (p.out2 && (reinterpret_cast<uintptr_t>(p.out2) & 7)) ||
            (p.out3 && (reinterpret_cast<uintptr_t>(p.out3) & 7)))
            return false;
    }
    return true;
}

// pair_request 2 = clusters of two pairs that share their A tiles by multicast (the caller checks `gemm2_quad_ok`).
bool gemm2_quad_ok(const GemmParams& p, int bn_tile) {
    const int n_tiles = (p.N + bn_tile - 1) / bn_tile;
    return p.mode == MOBI_EPI_PLAIN && bn_tile != 64 && n_tiles % 2 == 0 && p.batch <= 1 && !p.a_mn && !p.b_mn;
}

bool gemm2_pair_wanted(const GemmParams& p, int bn_tile, int pair_request) {
    if (pair_request < 0 || p.a_mn || p.b_mn) return false;
    if (pair_request > 0) return true;
    // auto: long-K tensor-bound problems with at least one full wave of 256-row tiles (74 CTA pairs on 148 SMs)
    const long long m2 = (p.M + 2 * BM - 1) / (2 * BM), nt = (p.N + bn_tile - 1) / bn_tile;
    const long long tiles = m2 * nt * (p.batch > 1 ? p.batch : 1);
    return p.num_k_blocks >= 16 && tiles >= sm_count() / 2;
}

int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const GemmParams& p, int bn_tile,
                 cudaStream_t stream) {
    switch (bn_tile) {
        case 64: return launch_gemm2_t<64>(tmA, tmB, tmR, p, stream);
        case 128: return launch_gemm2_t<128>(tmA, tmB, tmR, p, stream);
        case 160: return launch_gemm2_t<160>(tmA, tmB, tmR, p, stream);
        case 256: return launch_gemm2_t<256>(tmA, tmB, tmR, p, stream);
        default: MOBI_CHECK(false, "mobi_gemm: unsupported tile_n %d", bn_tile);
    }
    return 0;
}

}  // namespace mobi
