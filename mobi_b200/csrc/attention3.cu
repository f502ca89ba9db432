// Fused attention, third generation (head_dim <= 128, V in its natural [keys, d] layout).
//
// Softmax machinery (thread = query row, whole S row in registers from one tcgen05.ld pass,
// stale running maximum, packed-fp32 arithmetic, a share of the exponentials on the FMA pipe), plus:
//   * NT query tiles of 128 rows per CTA (NT = 4 for head_dim <= 64: ONE CTA per SM with 16 softmax warps, all four
//     tiles share one K/V ring, so K/V are fetched once per 512 query rows and the ring is 5 blocks deep: the TMA
//     round trip is off the critical path); every tile has its own MMA-issuer warp;
//   * V is consumed as an MN-major B operand of the PV product (tcgen05 instruction-descriptor bit 16, shared-memory
//     descriptor with LBO = stride between 64-wide d-chunks, SBO = 1024 B between 8-key groups), so the QKV
//     projection no longer has to store a transposed copy of V with 2-byte scattered writes.
// TMEM: S_t at columns t*BKV, O_t at NT*BKV + t*O_STRIDE (NT = 4: 4*64 + 4*64 = 512 columns).
// Warp roles: warps [0, 4NT) softmax (tile = warp / 4), warp 4NT TMA producer, warps 4NT+1+t MMA issuer of tile t
// (warp 4NT+1 also owns the TMEM allocation); the role warps donate registers to the softmax warps (setmaxnreg).
#include "../../include/mobi_b200.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int A3_CHUNK = 128 * 128;  // bytes of a 128-row x 64-col bf16 chunk

__device__ __forceinline__ float a3_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for a pair on the FMA pipe (Cody-Waite split + degree-3 polynomial)
__device__ __forceinline__ void a3_ex2_poly2(uint64_t x, float& e0, float& e1) {
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

template <int NT>
struct A3Cfg {
    static constexpr int ROLE_WARPS = NT <= 3 ? 4 : 8;
    static constexpr int WARPS = 4 * NT + ROLE_WARPS;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int SOFTMAX_REGS = NT == 4 ? 96 : 208;
};

constexpr uint32_t A3_ROLE_SLEEP_NS = 64;  // probe interval of the producer / issuer warps (see mbar_wait_backoff)

template <int NT, int BKV, int KVS, int POLY>
__global__ void __launch_bounds__(A3Cfg<NT>::THREADS, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using Cfg = A3Cfg<NT>;
    constexpr int KCH = BKV / 64;                      // 64-key chunks per block (P tiles)
    constexpr int K_CHUNK = BKV * 128;                 // bytes of a BKV-row x 64-col bf16 chunk of K or V
    constexpr int O_STRIDE = NT == 4 ? 64 : 128;       // TMEM columns reserved per O accumulator
    constexpr int O_BASE = NT * BKV;
    constexpr int BASE = 4 * NT;                       // first role warp
    static_assert(NT * (BKV + O_STRIDE) <= 512, "TMEM budget");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int q_tile_bytes = p.nch * A3_CHUNK;
    const int kv_bytes = p.nch * K_CHUNK;              // K and V blocks have the same footprint
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + NT * q_tile_bytes;
    uint8_t* sV = sK + KVS * kv_bytes;
    uint8_t* sP = sV + KVS * kv_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NT * KCH * A3_CHUNK);
    uint64_t* q_full = bars;
    uint64_t* s_full = bars + 1;
    uint64_t* s_free = s_full + NT;
    uint64_t* p_full = s_free + NT;
    uint64_t* pv_done = p_full + NT;
    uint64_t* kv_full = pv_done + NT;
    uint64_t* kv_empty = kv_full + KVS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + KVS);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (128 * NT);
    const int bh = blockIdx.y;
    const int nblk = (p.tk + BKV - 1) / BKV;
    const int ntiles = min(NT, (p.tq - q0 + 127) / 128);

    if (warp == BASE) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int i = 0; i < KVS; ++i) {
                mbar_init(&kv_full[i], 1);
                mbar_init(&kv_empty[i], ntiles);  // one tcgen05.commit per tile issuer
            }
            for (int i = 0; i < NT; ++i) {
                mbar_init(&s_full[i], 1);
                mbar_init(&s_free[i], 128);
                mbar_init(&p_full[i], 128);
                mbar_init(&pv_done[i], 1);
            }
            fence_barrier_init();
        }
    } else if (warp == BASE + 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= BASE) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");  // donate registers to the softmax warpgroups
        if (warp == BASE) {
            if (elect_one()) {
                // ---------------- TMA producer
                mbar_arrive_expect_tx(q_full, ntiles * q_tile_bytes);
                for (int t = 0; t < ntiles; ++t)
                    for (int c = 0; c < p.nch; ++c)
                        tma_load_3d(sQ + t * q_tile_bytes + c * A3_CHUNK, &tmQ, q_full, c * 64, q0 + t * 128, bh);
                for (int j = 0; j < nblk; ++j) {
                    const int s = j % KVS;
                    mbar_wait_backoff(&kv_empty[s], ((j / KVS) & 1) ^ 1, A3_ROLE_SLEEP_NS);
                    mbar_arrive_expect_tx(&kv_full[s], 2 * kv_bytes);
                    for (int c = 0; c < p.nch; ++c) {
                        tma_load_3d(sK + s * kv_bytes + c * K_CHUNK, &tmK, &kv_full[s], c * 64, j * BKV, bh);
                        tma_load_3d(sV + s * kv_bytes + c * K_CHUNK, &tmV, &kv_full[s], c * 64, j * BKV, bh);
                    }
                }
            }
        } else if (warp - (BASE + 1) < ntiles) {
            if (elect_one()) {
                // ---------------- MMA issuer of tile t: events of one tile arrive in program order (s_free(j) before
                // p_full(j)), so plain blocking waits never stall another tile
                const int t = warp - (BASE + 1);
                const uint32_t idesc_s = make_idesc_bf16(128, BKV);
                const uint32_t idesc_o = make_idesc_bf16(128, p.dn) | (1u << 16);  // B operand (V) is MN-major
                auto issue_S = [&](int j) {  // S_t(j) = Q_t K(j)^T
                    const int s = j % KVS;
                    const uint32_t d_tmem = tmem_base + t * BKV;
                    for (int k = 0; k < p.dk16; ++k) {
                        const uint64_t a =
                            make_kmajor_sw128_desc(smem_u32(sQ + t * q_tile_bytes + (k >> 2) * A3_CHUNK)) + 2 * (k & 3);
                        const uint64_t b =
                            make_kmajor_sw128_desc(smem_u32(sK + s * kv_bytes + (k >> 2) * K_CHUNK)) + 2 * (k & 3);
                        umma_bf16_ss(d_tmem, a, b, idesc_s, k != 0 ? 1u : 0u);
                    }
                    umma_commit(&s_full[t]);
                };
                mbar_wait_backoff(q_full, 0, A3_ROLE_SLEEP_NS);
                // staggered start: tile t begins once tile t-1 has pulled its first score tile into registers, so the
                // warpgroups sit in different phases (TMEM load / exponentials / P store) at any time
                if (t > 0) mbar_wait_backoff(&s_free[t - 1], 0, A3_ROLE_SLEEP_NS);
                mbar_wait_backoff(&kv_full[0], 0, A3_ROLE_SLEEP_NS);
                tc_fence_after();
                issue_S(0);
                for (int j = 0; j < nblk; ++j) {
                    const int s = j % KVS;
                    const uint32_t ph = j & 1;
                    if (j + 1 < nblk) {
                        mbar_wait_backoff(&kv_full[(j + 1) % KVS], ((j + 1) / KVS) & 1, A3_ROLE_SLEEP_NS);
                        mbar_wait_backoff(&s_free[t], ph, A3_ROLE_SLEEP_NS);
                        tc_fence_after();
                        issue_S(j + 1);
                    }
                    mbar_wait_backoff(&p_full[t], ph, A3_ROLE_SLEEP_NS);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + O_BASE + t * O_STRIDE;
                    for (int k = 0; k < BKV / 16; ++k) {  // 16 keys per MMA
                        const uint64_t a =
                            make_kmajor_sw128_desc(smem_u32(sP + (t * KCH + (k >> 2)) * A3_CHUNK)) + 2 * (k & 3);
                        const uint64_t b = make_mnmajor_sw128_desc(smem_u32(sV + s * kv_bytes + k * 2048), K_CHUNK);
                        umma_bf16_ss(d_tmem, a, b, idesc_o, (j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t]);
                    umma_commit(&kv_empty[s]);  // this tile is done with K(j) / V(j)
                }
            }
        }
    } else {
        if constexpr (Cfg::SOFTMAX_REGS == 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ---------------- softmax / correction / epilogue of tile t: thread <-> query row
        const int t = warp >> 2;
        if (t < ntiles) {
            const int lg = warp & 3;
            const int row = lg * 32 + lane;
            const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
            const uint32_t tS = tmem_base + t * BKV + lane_addr;
            const uint32_t tO = tmem_base + O_BASE + t * O_STRIDE + lane_addr;
            uint8_t* prow = sP + t * KCH * A3_CHUNK + (row >> 3) * 1024 + (row & 7) * 128;
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
                        mbar_wait(&pv_done[t], ph ^ 1);  // PV(j-1) complete before O is rescaled in place
                        tc_fence_after();
                        const float m_new = fmaxf(m_used, mx);
                        const float alpha = a3_ex2(m_used - m_new);
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
                    constexpr int kPolySlots[9] = {0x00, 0x08, 0x22, 0x2a, 0xaa, 0xab, 0xbb, 0xbf, 0xff};
                    if ((kPolySlots[POLY] >> ((c >> 1) & 7)) & 1) {
                        a3_ex2_poly2(x, e0, e1);
                    } else {
                        float x0, x1;
                        unpack2(x, x0, x1);
                        e0 = a3_ex2(x0);
                        e1 = a3_ex2(x1);
                    }
                    lsum2 = add2(lsum2, pack2(e0, e1));
                    sr[c >> 1] = pack_bf16x2(e0, e1);
                }
                float lsum, lsum_hi;
                unpack2(lsum2, lsum, lsum_hi);
                l += lsum + lsum_hi;
                if (j > 0) mbar_wait(&pv_done[t], ph ^ 1);  // P buffer free (almost always already true)
#pragma unroll
                for (int c = 0; c < BKV; c += 8) {
                    const int unit = (c & 63) >> 3;  // 16-byte unit inside the 128-byte row
                    *reinterpret_cast<uint4*>(prow + (c >> 6) * A3_CHUNK + ((unit ^ (row & 7)) << 4)) =
                        make_uint4(sr[(c >> 1)], sr[(c >> 1) + 1], sr[(c >> 1) + 2], sr[(c >> 1) + 3]);
                }
                fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
                tc_fence_before();
                mbar_arrive(&p_full[t]);
            }
            // ---------------- epilogue: O / l -> out[b, t, h*d + :]
            mbar_wait(&pv_done[t], (nblk - 1) & 1);
            tc_fence_after();
            const float inv_l = 1.0f / l;
            const int tq_row = q0 + t * 128 + row;
            if (p.lse != nullptr && tq_row < p.tq) p.lse[(long long)bh * p.tq + tq_row] = m_used + log2f(l);
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
    if (warp == BASE + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int NT, int BKV, int KVS, int POLY>
static int launch_attention3(const mobi_attn_args* a, AttnParams p, cudaStream_t stream) {
    const int d = a->head_dim;
    const long long BH = (long long)a->batch * a->heads;
    const long long smem = (long long)NT * p.nch * A3_CHUNK + 2ll * KVS * p.nch * BKV * 128 +
                           (long long)NT * (BKV / 64) * A3_CHUNK + 512 + 1024;
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
        if (make_tensor_map_bf16(&tmV, a->vt, 3, dims, strides, box)) return 1;  // V: [BH, Tk, d] like K
    }
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(attention3_kernel<NT, BKV, KVS, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)limit));
        configured = true;
    }
    dim3 grid((a->tq + 128 * NT - 1) / (128 * NT), (unsigned)BH, 1);
    attention3_kernel<NT, BKV, KVS, POLY><<<grid, A3Cfg<NT>::THREADS, smem, stream>>>(tmQ, tmK, tmV, p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

int attention3_dispatch(const mobi_attn_args* a, const AttnParams& p, cudaStream_t stream) {
    if (a->head_dim <= 64) {
        if ((a->kernel & 15) == 4) return launch_attention3<2, 64, 3, 2>(a, p, stream);  // tuning hook: 2 tiles
        return launch_attention3<4, 64, 5, 2>(a, p, stream);
    }
    return launch_attention3<2, 64, 3, 2>(a, p, stream);
}

}  // namespace mobi
