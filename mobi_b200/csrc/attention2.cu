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
// issuer + TMEM owner, warps 10-11 idle (they only donate registers: setmaxnreg).
#include "../../include/mobi_b200.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int A2_CHUNK = 128 * 128;  // bytes of a 128-row x 64-col bf16 chunk

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// MINB = 2: two CTAs per SM (BKV = 64, head_dim <= 64): four softmax warps per SM sub-partition instead of two.
template <int BKV, int MINB>
__global__ void __launch_bounds__(384, MINB)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    constexpr int KCH = BKV / 64;            // 64-key chunks per block (P / V tiles)
    constexpr int O_STRIDE = MINB == 2 ? 64 : 128;             // TMEM columns reserved per O accumulator
    constexpr int TMEM_COLS = MINB == 2 ? 256 : 512;           // S: 2 x BKV columns, O: 2 x O_STRIDE columns
    constexpr int O_BASE = 2 * BKV;
    constexpr int K_CHUNK = BKV * 128;       // bytes of a BKV-row x 64-col bf16 chunk of K
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int q_tile_bytes = p.nch * A2_CHUNK;
    const int k_bytes = p.nch * K_CHUNK;
    const int v_chunk = p.dn * 128;
    const int v_bytes = KCH * v_chunk;
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 2 * q_tile_bytes;
    uint8_t* sV = sK + 2 * k_bytes;
    uint8_t* sP = sV + 2 * v_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * KCH * A2_CHUNK);
    uint64_t* q_full = bars;         // 1
    uint64_t* kv_full = bars + 1;    // 2
    uint64_t* kv_empty = bars + 3;   // 2
    uint64_t* s_full = bars + 5;     // 2 (per tile)
    uint64_t* s_free = bars + 7;     // 2
    uint64_t* p_full = bars + 9;     // 2
    uint64_t* pv_done = bars + 11;   // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

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
            for (int i = 0; i < 2; ++i) {
                mbar_init(&kv_full[i], 1);
                mbar_init(&kv_empty[i], 1);
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
                    const int s = j & 1;
                    mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&kv_full[s], k_bytes + v_bytes);
                    for (int c = 0; c < p.nch; ++c)
                        tma_load_3d(sK + s * k_bytes + c * K_CHUNK, &tmK, &kv_full[s], c * 64, j * BKV, bh);
                    for (int c = 0; c < KCH; ++c)
                        tma_load_3d(sV + s * v_bytes + c * v_chunk, &tmV, &kv_full[s], j * BKV + c * 64, 0, bh);
                }
            }
        } else if (warp == 9) {
            if (elect_one()) {
                // ---------------- MMA issuer
                const uint32_t idesc_s = make_idesc_bf16(128, BKV);
                const uint32_t idesc_o = make_idesc_bf16(128, p.dn);
                auto issue_S = [&](int t, int j) {  // S_t(j) = Q_t K(j)^T
                    const int s = j & 1;
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
                auto issue_PV = [&](int t, int j) {  // O_t += P_t(j) V(j)
                    const int s = j & 1;
                    const uint32_t d_tmem = tmem_base + O_BASE + t * O_STRIDE;
                    for (int k = 0; k < BKV / 16; ++k) {
                        const uint64_t a =
                            make_kmajor_sw128_desc(smem_u32(sP + (t * KCH + (k >> 2)) * A2_CHUNK)) + 2 * (k & 3);
                        const uint64_t b =
                            make_kmajor_sw128_desc(smem_u32(sV + s * v_bytes + (k >> 2) * v_chunk)) + 2 * (k & 3);
                        umma_bf16_ss(d_tmem, a, b, idesc_o, (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t]);
                };
                // Event loop: the two tiles advance independently (no head-of-line blocking between them).
                //   S_t(j+1) is issued as soon as softmax t has pulled S_t(j) into registers (s_free) and K(j+1) landed;
                //   PV_t(j)  is issued as soon as P_t(j) is in shared memory (p_full);
                //   a K/V stage is released once both tiles have issued their PV on it.
                // Tile 1 starts only after tile 0 has read its first score tile: the half-block phase offset keeps one
                // warpgroup in its exponential phase while the other one waits for TMEM loads.
                int js[2] = {0, 0};  // next score block to issue per tile
                int jp[2] = {0, 0};  // next PV block to issue per tile
                int released = 0;    // K/V blocks handed back to the producer
                mbar_wait(q_full, 0);
                const long long t_start = clock64();
                for (;;) {
                    bool done = true, progress = false;
                    for (int t = 0; t < ntiles; ++t) {
                        if (jp[t] < nblk) done = false;
                        // next score tile
                        if (js[t] < nblk) {
                            const int j = js[t];
                            // (probes are non-blocking and each barrier is at most one phase ahead of the one probed)
                            bool ok = mbar_test_wait(&kv_full[j & 1], (j >> 1) & 1);
                            if (ok && j > 0) ok = mbar_test_wait(&s_free[t], (j - 1) & 1);
                            if (ok && j == 0 && t == 1) ok = (js[0] >= 2) || (jp[0] >= 1);  // phase offset
                            if (ok) {
                                tc_fence_after();
                                issue_S(t, j);
                                js[t] = j + 1;
                                progress = true;
                            }
                        }
                        // next value product
                        if (jp[t] < nblk && jp[t] < js[t]) {
                            const int j = jp[t];
                            if (mbar_test_wait(&p_full[t], j & 1)) {
                                tc_fence_after();
                                issue_PV(t, j);
                                jp[t] = j + 1;
                                progress = true;
                                const int lo = ntiles == 2 ? min(jp[0], jp[1]) : jp[0];
                                while (released < lo) {
                                    umma_commit(&kv_empty[released & 1]);
                                    ++released;
                                }
                            }
                        }
                    }
                    if (done) break;
                    if (!progress && clock64() - t_start > MOBI_WAIT_LIMIT_CYCLES) asm volatile("trap;");
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
                float lsum = 0.f;
#pragma unroll
                for (int c = 0; c < BKV; c += 8) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) e[i] = ex2_approx(__uint_as_float(sr[c + i]) - m_used);
                    lsum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
#pragma unroll
                    for (int i = 0; i < 4; ++i) sr[(c >> 1) + i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
                }
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

template <int BKV, int MINB>
static int launch_attention2(const mobi_attn_args* a, AttnParams p, cudaStream_t stream) {
    const int d = a->head_dim;
    const long long BH = (long long)a->batch * a->heads;
    const long long smem = 2ll * p.nch * A2_CHUNK + 2ll * p.nch * BKV * 128 + 2ll * (BKV / 64) * p.dn * 128 +
                           2ll * (BKV / 64) * A2_CHUNK + 256 + 1024;
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
        MOBI_CUDA(cudaFuncSetAttribute(attention2_kernel<BKV, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        configured = true;
    }
    dim3 grid((a->tq + 255) / 256, (unsigned)BH, 1);
    attention2_kernel<BKV, MINB><<<grid, 384, smem, stream>>>(tmQ, tmK, tmV, p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

int attention2_dispatch(const mobi_attn_args* a, const AttnParams& p, cudaStream_t stream) {
    if (a->head_dim <= 64 && a->kernel != 2) return launch_attention2<64, 2>(a, p, stream);
    if (a->head_dim <= 64) return launch_attention2<128, 1>(a, p, stream);
    return launch_attention2<64, 1>(a, p, stream);
}

}  // namespace mobi
