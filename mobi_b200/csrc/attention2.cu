// Flash-style fused attention, second generation (head_dim <= 128): two query tiles per CTA, two softmax
// warpgroups, MUFU-bound by design.
//
// Why: at MObI's head dims (40 and 80) attention is NOT tensor-core bound.  Per 128x128 score tile the tensor
// pipe needs 4*d*128*128 / 8192 FLOP/clk = 320 clk (d = 40) while the 16384 exponentials need 1024 clk of the
// SM's 16/clk MUFU.  So the design goal is to keep the MUFU busy every cycle:
//   * a CTA owns 256 query rows of one (batch, head) as two 128-row tiles; each tile has its own softmax
//     warpgroup (thread == query row == TMEM lane: no shuffles), its own S accumulator (128 TMEM columns), its own
//     O accumulator and its own P staging buffer, so the two warpgroups never wait for each other and the four
//     SM sub-partitions always have two resident softmax warps to interleave;
//   * each softmax thread pulls its whole S row into registers with ONE pass of tcgen05.ld, releases the S
//     columns at once (s_free), and the MMA warp issues the NEXT QK^T for that tile while the exponentials of the
//     current one are still running - the tensor pipe is always a full tile ahead;
//   * running maximum is stale-by-at-most-2^8 (rescale of O in TMEM only when a row maximum grows by > 8 in
//     log2 units), the softmax scale and log2(e) are folded into W_q, so the inner loop per element is
//     FADD, MUFU.EX2, FADD, half an F2FP and 1/8 of a 16-byte st.shared.
// K/V blocks stream through a 2-stage TMA ring; P goes to 128B-swizzled shared memory as the A operand of
// O += P V; scores and probabilities never touch HBM.
//
// Warp roles (384 threads): warps 0-3 softmax tile 0, warps 4-7 softmax tile 1, warp 8 TMA producer, warp 9 MMA
// issuer of tile 0 + TMEM owner, warp 10 MMA issuer of tile 1, warp 11 idle (warps 8-11 donate registers: setmaxnreg).
#include "../../include/mobi_b200.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int A2_CHUNK = 128 * 128;  // bytes of a 128-row x 64-col bf16 chunk
constexpr int A2_KV_STAGES = 3;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for a pair of arguments on the FMA pipe instead of the MUFU (Cody-Waite split + degree-3 minimax polynomial,
// relative error ~1e-4, far below the bf16 resolution of P): x = n + f with n = floor(x) taken from the mantissa of
// x + 1.5*2^23 (round toward -inf), 2^f from the polynomial, and n added straight into the exponent field.
// 2 FMNMX + 3 FADD2 + 3 FFMA2 + 2 LEA per pair.  Arguments are <= 8 here (stale running maximum) and clamped at -127.
__device__ __forceinline__ void ex2_poly2(uint64_t x, float& e0, float& e1) {
    float x0, x1;
    unpack2(x, x0, x1);
    const uint64_t magic = pack2(12582912.f, 12582912.f);
    const uint64_t xc = pack2(fmaxf(x0, -127.f), fmaxf(x1, -127.f));
    const uint64_t t = add2_rm(xc, magic);
    const uint64_t f = sub2(xc, sub2(t, magic));
    uint64_t pl = fma2(pack2(0.077119089663028717f, 0.077119089663028717f), f,
                       pack2(0.227564394474029541f, 0.227564394474029541f));
    pl = fma2(pl, f, pack2(0.695146143436431885f, 0.695146143436431885f));
    pl = fma2(pl, f, pack2(1.f, 1.f));
    float p0, p1, t0, t1;
    unpack2(pl, p0, p1);
    unpack2(t, t0, t1);
    e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// MINB = 2: two CTAs per SM (BKV = 64, head_dim <= 64): four softmax warps per SM sub-partition instead of two.
// POLY = how many of every 8 consecutive pairs of exponentials go to the FMA pipe (the rest use MUFU.EX2): the MUFU
// (16/clk/SM) is the bottleneck of this kernel, so a fraction of the work is moved to where there are spare issue slots.
template <int BKV, int MINB, int POLY>
__global__ void __launch_bounds__(384, MINB)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    constexpr int KCH = BKV / 64;            // 64-key chunks per block (P / V tiles)
    constexpr int O_STRIDE = MINB == 2 ? 64 : 128;             // TMEM columns reserved per O accumulator
    constexpr int TMEM_COLS = MINB == 2 ? 256 : 512;           // S: 2 x BKV columns, O: 2 x O_STRIDE columns
    constexpr int O_BASE = 2 * BKV;
    constexpr int K_CHUNK = BKV * 128;       // bytes of a BKV-row x 64-col bf16 chunk of K
    constexpr int KVS = A2_KV_STAGES;        // K/V ring depth: the TMA round trip (~1 us) spans more than one block
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int q_tile_bytes = p.nch * A2_CHUNK;
    const int k_bytes = p.nch * K_CHUNK;
    const int v_chunk = p.dn * 128;
    const int v_bytes = KCH * v_chunk;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 2 * q_tile_bytes;
    uint8_t* sV = sK + KVS * k_bytes;
    uint8_t* sP = sV + KVS * v_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * KCH * A2_CHUNK);
    uint64_t* q_full = bars;         // 1
    uint64_t* s_full = bars + 1;     // 2 (per tile)
    uint64_t* s_free = bars + 3;     // 2
    uint64_t* p_full = bars + 5;     // 2
    uint64_t* pv_done = bars + 7;    // 2
    uint64_t* kv_full = bars + 9;    // KVS
    uint64_t* kv_empty = kv_full + KVS;  // KVS
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + KVS);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256;
    const int bh = blockIdx.y;
    const int nblk = (p.tk + BKV - 1) / BKV;
    const int ntiles = (q0 + 128 < p.tq) ? 2 : 1;

    if (warp == 8) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int i = 0; i < KVS; ++i) {
                mbar_init(&kv_full[i], 1);
                mbar_init(&kv_empty[i], ntiles);  // one tcgen05.commit per tile issuer
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s_full[i], 1);
                mbar_init(&s_free[i], 128);
                mbar_init(&p_full[i], 128);
                mbar_init(&pv_done[i], 1);
            }
            fence_barrier_init();
        }
    } else if (warp == 9) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");  // donate registers to the softmax warpgroups
        if (warp == 8) {
            if (elect_one()) {
                // ---------------- TMA producer
                mbar_arrive_expect_tx(q_full, ntiles * q_tile_bytes);
                for (int t = 0; t < ntiles; ++t)
                    for (int c = 0; c < p.nch; ++c)
                        tma_load_3d(sQ + t * q_tile_bytes + c * A2_CHUNK, &tmQ, q_full, c * 64, q0 + t * 128, bh);
                for (int j = 0; j < nblk; ++j) {
                    const int s = j % KVS;
                    mbar_wait(&kv_empty[s], ((j / KVS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&kv_full[s], k_bytes + v_bytes);
                    for (int c = 0; c < p.nch; ++c)
                        tma_load_3d(sK + s * k_bytes + c * K_CHUNK, &tmK, &kv_full[s], c * 64, j * BKV, bh);
                    for (int c = 0; c < KCH; ++c)
                        tma_load_3d(sV + s * v_bytes + c * v_chunk, &tmV, &kv_full[s], j * BKV + c * 64, 0, bh);
                }
            }
        } else if (warp - 9 < ntiles) {
            if (elect_one()) {
                // ---------------- MMA issuer of tile t (warp 9: tile 0, warp 10: tile 1).  Each tile has its own issuing
                // thread, so the two tiles never block each other; inside a tile the events arrive in program order
                // (s_free(j) always precedes p_full(j)), so plain blocking waits have no head-of-line stalls.
                const int t = warp - 9;
                const uint32_t idesc_s = make_idesc_bf16(128, BKV);
                const uint32_t idesc_o = make_idesc_bf16(128, p.dn);
                auto issue_S = [&](int j) {  // S_t(j) = Q_t K(j)^T
                    const int s = j % KVS;
                    const uint32_t d_tmem = tmem_base + t * BKV;
                    for (int k = 0; k < p.dk16; ++k) {
                        const uint64_t a =
                            make_kmajor_sw128_desc(smem_u32(sQ + t * q_tile_bytes + (k >> 2) * A2_CHUNK)) + 2 * (k & 3);
                        const uint64_t b =
                            make_kmajor_sw128_desc(smem_u32(sK + s * k_bytes + (k >> 2) * K_CHUNK)) + 2 * (k & 3);
                        umma_bf16_ss(d_tmem, a, b, idesc_s, k != 0 ? 1u : 0u);
                    }
                    umma_commit(&s_full[t]);
                };
                mbar_wait(q_full, 0);
                // tile 1 starts once tile 0 has pulled its first score tile into registers: the half-block phase offset
                // keeps one warpgroup in its exponential phase while the other waits for TMEM loads
                if (t == 1) mbar_wait(&s_free[0], 0);
                mbar_wait(&kv_full[0], 0);
                tc_fence_after();
                issue_S(0);
                for (int j = 0; j < nblk; ++j) {
                    const int s = j % KVS;
                    const uint32_t ph = j & 1;
                    if (j + 1 < nblk) {
                        // next score tile as soon as the softmax threads have S(j) in registers and K(j+1) has landed
                        mbar_wait(&kv_full[(j + 1) % KVS], ((j + 1) / KVS) & 1);
                        mbar_wait(&s_free[t], ph);
                        tc_fence_after();
                        issue_S(j + 1);
                    }
                    mbar_wait(&p_full[t], ph);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + O_BASE + t * O_STRIDE;
                    for (int k = 0; k < BKV / 16; ++k) {
                        const uint64_t a =
                            make_kmajor_sw128_desc(smem_u32(sP + (t * KCH + (k >> 2)) * A2_CHUNK)) + 2 * (k & 3);
                        const uint64_t b =
                            make_kmajor_sw128_desc(smem_u32(sV + s * v_bytes + (k >> 2) * v_chunk)) + 2 * (k & 3);
                        umma_bf16_ss(d_tmem, a, b, idesc_o, (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t]);
                    umma_commit(&kv_empty[s]);  // this tile is done with K(j) / V(j)
                }
            }
        }
    } else {
        if constexpr (MINB == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ---------------- softmax / correction / epilogue of tile t: thread <-> query row
        const int t = warp >> 2;
        if (t < ntiles) {
            const int lg = warp & 3;
            const int row = lg * 32 + lane;
            const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
            const uint32_t tS = tmem_base + t * BKV + lane_addr;
            const uint32_t tO = tmem_base + O_BASE + t * O_STRIDE + lane_addr;
            uint8_t* prow = sP + t * KCH * A2_CHUNK + (row >> 3) * 1024 + (row & 7) * 128;
            float m_used = -INFINITY;
            float l = 0.f;
            for (int j = 0; j < nblk; ++j) {
                const uint32_t ph = j & 1;
                uint32_t sr[BKV];
                mbar_wait(&s_full[t], ph);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < BKV; c += 32) tmem_ld32(tS + c, reinterpret_cast<uint32_t(&)[32]>(sr[c]));
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[t]);  // S columns may be overwritten by the next QK^T
                const int valid = p.tk - j * BKV;  // keys of this block that exist (>= BKV except in the last block)
                if (valid < BKV) {
#pragma unroll
                    for (int i = 0; i < BKV; ++i)
                        if (i >= valid) sr[i] = 0xff800000u;  // -inf
                }
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < BKV; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
                if (j == 0) {
                    m_used = mx;
                } else {
                    const bool need = mx > m_used + 8.0f;
                    if (__any_sync(0xffffffffu, need)) {
                        // PV(j-1) must be complete before O is rescaled in place
                        mbar_wait(&pv_done[t], ph ^ 1);
                        tc_fence_after();
                        const float m_new = fmaxf(m_used, mx);
                        const float alpha = ex2_approx(m_used - m_new);
                        l *= alpha;
                        m_used = m_new;
#pragma unroll 1
                        for (int c = 0; c < p.dn; c += 16) {
                            uint32_t r[16];
                            tmem_ld16(tO + c, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            tmem_st16(tO + c, r);
                        }
                        tmem_st_wait();
                    }
                }
                // probabilities -> packed bf16 in registers (in place: sr[c/2] <- pack(p[c], p[c+1]))
                const uint64_t m2 = pack2(m_used, m_used);
                uint64_t lsum2 = pack2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < BKV; c += 2) {
                    const uint64_t x = sub2(pack2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), m2);
                    float e0, e1;
                    // pair index within its group of 8 pairs decides the pipe (spread so the two pipes interleave)
                    constexpr int kPolySlots[9] = {0x00, 0x08, 0x22, 0x2a, 0xaa, 0xab, 0xbb, 0xbf, 0xff};
                    if ((kPolySlots[POLY] >> ((c >> 1) & 7)) & 1) {
                        ex2_poly2(x, e0, e1);
                    } else {
                        float x0, x1;
                        unpack2(x, x0, x1);
                        e0 = ex2_approx(x0);
                        e1 = ex2_approx(x1);
                    }
                    lsum2 = add2(lsum2, pack2(e0, e1));
                    sr[c >> 1] = pack_bf16x2(e0, e1);
                }
                float lsum, lsum_hi;
                unpack2(lsum2, lsum, lsum_hi);
                lsum += lsum_hi;
                // the P buffer is free once PV(j-1) has completed (it almost always has by now)
                if (j > 0) mbar_wait(&pv_done[t], ph ^ 1);
                // swizzled K-major tile (row = query, 64 keys per 128-byte row chunk)
#pragma unroll
                for (int c = 0; c < BKV; c += 8) {
                    const int unit = (c & 63) >> 3;  // 16-byte unit inside the 128-byte row
                    *reinterpret_cast<uint4*>(prow + (c >> 6) * A2_CHUNK + ((unit ^ (row & 7)) << 4)) =
                        make_uint4(sr[(c >> 1)], sr[(c >> 1) + 1], sr[(c >> 1) + 2], sr[(c >> 1) + 3]);
                }
                l += lsum;
                fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                tc_fence_before();
                mbar_arrive(&p_full[t]);
            }
            // ---------------- epilogue: O / l -> out[b, t, h*d + :]
            mbar_wait(&pv_done[t], (nblk - 1) & 1);
            tc_fence_after();
            const float inv_l = 1.0f / l;
            const int tq_row = q0 + t * 128 + row;
            const int b = bh / p.heads, h = bh - b * p.heads;
            __nv_bfloat16* orow = p.out + ((long long)b * p.tq + tq_row) * p.ld_out + h * p.head_dim;
#pragma unroll 1
            for (int c = 0; c < p.dn; c += 16) {
                uint32_t r[16];
                tmem_ld16(tO + c, r);
                tmem_ld_wait();
                if (tq_row < p.tq) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (c + u * 8 + 8 <= p.head_dim) {
                            uint4 pk;
                            pk.x = pack_bf16x2(__uint_as_float(r[u * 8 + 0]) * inv_l, __uint_as_float(r[u * 8 + 1]) * inv_l);
                            pk.y = pack_bf16x2(__uint_as_float(r[u * 8 + 2]) * inv_l, __uint_as_float(r[u * 8 + 3]) * inv_l);
                            pk.z = pack_bf16x2(__uint_as_float(r[u * 8 + 4]) * inv_l, __uint_as_float(r[u * 8 + 5]) * inv_l);
                            pk.w = pack_bf16x2(__uint_as_float(r[u * 8 + 6]) * inv_l, __uint_as_float(r[u * 8 + 7]) * inv_l);
                            *reinterpret_cast<uint4*>(orow + c + u * 8) = pk;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BKV, int MINB, int POLY>
static int launch_attention2(const mobi_attn_args* a, AttnParams p, cudaStream_t stream) {
    const int d = a->head_dim;
    const long long BH = (long long)a->batch * a->heads;
    const long long smem = 2ll * p.nch * A2_CHUNK + (long long)A2_KV_STAGES * p.nch * BKV * 128 +
                           (long long)A2_KV_STAGES * (BKV / 64) * p.dn * 128 + 2ll * (BKV / 64) * A2_CHUNK + 256 + 1024;
    const long long limit = 227 * 1024;
    MOBI_CHECK(smem <= limit, "mobi_attention: head_dim=%d needs %lld bytes of shared memory", d, smem);
    CUtensorMap tmQ, tmK, tmV;
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)a->tq, (uint64_t)BH};
        uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)a->tq * d * 2};
        uint32_t box[3] = {64, 128, 1};
        if (make_tensor_map_bf16(&tmQ, a->q, 3, dims, strides, box)) return 1;
    }
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)a->tk, (uint64_t)BH};
        uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)a->tk * d * 2};
        uint32_t box[3] = {64, BKV, 1};
        if (make_tensor_map_bf16(&tmK, a->k, 3, dims, strides, box)) return 1;
    }
    {
        uint64_t dims[3] = {(uint64_t)a->tk, (uint64_t)d, (uint64_t)BH};
        uint64_t strides[2] = {(uint64_t)a->tk * 2, (uint64_t)a->tk * d * 2};
        uint32_t box[3] = {64, (uint32_t)p.dn, 1};
        if (make_tensor_map_bf16(&tmV, a->vt, 3, dims, strides, box)) return 1;
    }
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(attention2_kernel<BKV, MINB, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        configured = true;
    }
    dim3 grid((a->tq + 255) / 256, (unsigned)BH, 1);
    attention2_kernel<BKV, MINB, POLY><<<grid, 384, smem, stream>>>(tmQ, tmK, tmV, p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

int attention2_dispatch(const mobi_attn_args* a, const AttnParams& p, cudaStream_t stream) {
    const int poly = (a->kernel >> 4) & 15;  // tuning hook: 0 = default split, 1..8 = pairs of 8 on the FMA pipe, 9 = none
    if (a->head_dim <= 64 && (a->kernel & 15) != 2) {
        switch (poly) {
            case 9: return launch_attention2<64, 2, 0>(a, p, stream);
            case 2: return launch_attention2<64, 2, 2>(a, p, stream);
            case 4: return launch_attention2<64, 2, 4>(a, p, stream);
            case 3: return launch_attention2<64, 2, 3>(a, p, stream);
            default: return launch_attention2<64, 2, 2>(a, p, stream);
        }
    }
    if (a->head_dim <= 64) return launch_attention2<128, 1, 2>(a, p, stream);
    return launch_attention2<64, 1, 2>(a, p, stream);
}

}  // namespace mobi
