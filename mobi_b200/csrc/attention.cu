// Flash-style fused attention on tcgen05 for sm_100a.
//
// One CTA owns 128 query rows of one (batch, head).  Key/value blocks of 128 keys stream through shared
// memory by TMA; the score tile S = Q K^T (128x128 fp32) lives in TMEM (double buffered), each of the 128
// softmax threads owns one query row (TMEM lane == row, so no cross-thread reductions), writes the
// probabilities as a bf16 128B-swizzled K-major tile to shared memory, and the value product
// O += P V accumulates in TMEM.  The running maximum is only refreshed when it grows by more than 2^8
// (stale-max trick), so the O rescale in TMEM is rare.  Scores never touch HBM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = softmax /
// correction / epilogue.
#include "../../include/mobi_b200.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BKV = 128;  // keys per block
constexpr int ATT_TILE = 128 * 128;  // bytes of a 128-row x 64-col bf16 chunk

__host__ __device__ inline int att_v_tile_bytes(int dn) { return dn * 128; }

__global__ void __launch_bounds__(192, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int q_bytes = p.nch * ATT_TILE;
    const int k_bytes = p.nch * ATT_TILE;
    const int v_bytes = 2 * att_v_tile_bytes(p.dn);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + q_bytes;
    uint8_t* sV = sK + p.kv_stages * k_bytes;
    uint8_t* sP = sV + p.kv_stages * v_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + p.p_bufs * 2 * ATT_TILE);
    uint64_t* q_full = bars;          // 1
    uint64_t* kv_full = bars + 1;     // 2
    uint64_t* kv_empty = bars + 3;    // 2
    uint64_t* s_full = bars + 5;      // 2
    uint64_t* p_full = bars + 7;      // 2
    uint64_t* pv_done = bars + 9;     // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BM;
    const int bh = blockIdx.y;
    const int nblk = (p.tk + ATT_BKV - 1) / ATT_BKV;

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
            mbar_init(q_full, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&kv_full[i], 1);
                mbar_init(&kv_empty[i], 1);
                mbar_init(&s_full[i], 1);
                mbar_init(&p_full[i], 128);
                mbar_init(&pv_done[i], 1);
            }
            fence_barrier_init();
        }
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    pdl_trigger();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S0 = tmem_base;          // columns [0,128) and [128,256)
    const uint32_t tmem_O = tmem_base + 256;     // columns [256, 256+dn)

    if (warp == 0) {
        if (elect_one()) {
            // ---------------- TMA producer
            mbar_arrive_expect_tx(q_full, q_bytes);
            for (int c = 0; c < p.nch; ++c) tma_load_3d(sQ + c * ATT_TILE, &tmQ, q_full, c * 64, q0, bh);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % p.kv_stages;
                const uint32_t ph = (j / p.kv_stages) & 1;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], k_bytes + v_bytes);
                for (int c = 0; c < p.nch; ++c)
                    tma_load_3d(sK + s * k_bytes + c * ATT_TILE, &tmK, &kv_full[s], c * 64, j * ATT_BKV, bh);
                for (int c = 0; c < 2; ++c)
                    tma_load_3d(sV + s * v_bytes + c * att_v_tile_bytes(p.dn), &tmV, &kv_full[s],
                                j * ATT_BKV + c * 64, 0, bh);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ---------------- MMA issuer
            const uint32_t idesc_s = make_idesc_bf16(128, ATT_BKV);
            const uint32_t idesc_o = make_idesc_bf16(128, p.dn);
            auto issue_S = [&](int j) {
                const int s = j % p.kv_stages;
                mbar_wait(&kv_full[s], (j / p.kv_stages) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_S0 + (j & 1) * 128;
                for (int k = 0; k < p.dk16; ++k) {
                    const uint64_t a = make_kmajor_sw128_desc(smem_u32(sQ + (k >> 2) * ATT_TILE)) + 2 * (k & 3);
                    const uint64_t b =
                        make_kmajor_sw128_desc(smem_u32(sK + s * k_bytes + (k >> 2) * ATT_TILE)) + 2 * (k & 3);
                    umma_bf16_ss(d_tmem, a, b, idesc_s, k != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            for (int j = 0; j < p.kv_stages && j < nblk; ++j) issue_S(j);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % p.kv_stages;
                const int pb = j % p.p_bufs;
                mbar_wait(&p_full[pb], (j / p.p_bufs) & 1);
                tc_fence_after();
                for (int k = 0; k < ATT_BKV / 16; ++k) {
                    const uint64_t a =
                        make_kmajor_sw128_desc(smem_u32(sP + (pb * 2 + (k >> 2)) * ATT_TILE)) + 2 * (k & 3);
                    const uint64_t b = make_kmajor_sw128_desc(
                                           smem_u32(sV + s * v_bytes + (k >> 2) * att_v_tile_bytes(p.dn))) +
                                       2 * (k & 3);
                    umma_bf16_ss(tmem_O, a, b, idesc_o, (j | k) != 0 ? 1u : 0u);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(&pv_done[pb]);
                if (j + p.kv_stages < nblk) issue_S(j + p.kv_stages);
            }
        }
    } else {
        // ---------------- softmax / correction / epilogue: thread <-> query row
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
        float m_used = -INFINITY;
        float l = 0.f;
        for (int j = 0; j < nblk; ++j) {
            const uint32_t tS = tmem_S0 + (j & 1) * 128 + lane_addr;
            const int valid = min(ATT_BKV, p.tk - j * ATT_BKV);  // keys of this block that exist
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            // pass 1: block row maximum
            float m_blk = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < ATT_BKV; c += 32) {
                uint32_t r[32];
                __syncwarp();
                tmem_ld32(tS + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c + i < valid) m_blk = fmaxf(m_blk, __uint_as_float(r[i]));
            }
            if (j == 0) {
                m_used = m_blk;
            } else {
                const bool need = m_blk > m_used + 8.0f;
                if (__any_sync(0xffffffffu, need)) {
                    const float m_new = fmaxf(m_used, m_blk);
                    const float alpha = exp2f(m_used - m_new);
                    l *= alpha;
                    m_used = m_new;
                    // O must be complete (PV of block j-1) before it is rescaled in place
                    mbar_wait(&pv_done[(j - 1) % p.p_bufs], ((j - 1) / p.p_bufs) & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < p.dn; c += 16) {
                        uint32_t r[16];
                        __syncwarp();
                        tmem_ld16(tmem_O + lane_addr + c, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st16(tmem_O + lane_addr + c, r);
                    }
                    tmem_st_wait();
                }
            }
            // the P buffer is free once the PV product that last read it has completed
            const int pb = j % p.p_bufs;
            if (j >= p.p_bufs) mbar_wait(&pv_done[pb], ((j / p.p_bufs) - 1) & 1);
            // pass 2: probabilities -> bf16, swizzled K-major tile (row = query, 64 keys per chunk)
            uint8_t* prow = sP + pb * 2 * ATT_TILE + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll 1
            for (int c = 0; c < ATT_BKV; c += 32) {
                uint32_t r[32];
                __syncwarp();
                tmem_ld32(tS + c, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (c + i < valid) ? exp2f(__uint_as_float(r[i]) - m_used) : 0.f;
                    float p1 = (c + i + 1 < valid) ? exp2f(__uint_as_float(r[i + 1]) - m_used) : 0.f;
                    l += p0 + p1;
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                }
                uint8_t* chunk = prow + (c >> 6) * ATT_TILE;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int unit = ((c & 63) >> 3) + u;  // 16-byte unit inside the 128-byte row
                    *reinterpret_cast<uint4*>(chunk + ((unit ^ (row & 7)) << 4)) =
                        make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                }
            }
            fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            tc_fence_before();     // orders the TMEM reads/writes above before the MMA that follows the barrier
            mbar_arrive(&p_full[pb]);
        }
        // ---------------- epilogue: O / l -> out[b, t, h*d + :]
        mbar_wait(&pv_done[(nblk - 1) % p.p_bufs], ((nblk - 1) / p.p_bufs) & 1);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const int t = q0 + row;
        const int b = bh / p.heads, h = bh - b * p.heads;
        __nv_bfloat16* orow = p.out + ((long long)b * p.tq + t) * p.ld_out + h * p.head_dim;
#pragma unroll 1
        for (int c = 0; c < p.dn; c += 16) {
            uint32_t r[16];
            __syncwarp();
            tmem_ld16(tmem_O + lane_addr + c, r);
            tmem_ld_wait();
            if (t < p.tq) {
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
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace mobi

using namespace mobi;

extern "C" int mobi_attention(const mobi_attn_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->q && a->k && a->vt && a->out, "mobi_attention: null argument");
    const int d = a->head_dim;
    MOBI_CHECK(d % 8 == 0 && d >= 8 && d <= 192, "mobi_attention: head_dim=%d must be a multiple of 8 in [8,192]", d);
    MOBI_CHECK(a->tq > 0 && a->tk > 0 && (a->v_rowmajor || a->tk % 8 == 0),
               "mobi_attention: tq=%d tk=%d (tk must be a multiple of 8 for the transposed V layout)", a->tq, a->tk);
    MOBI_CHECK(a->ld_out % 8 == 0 && a->ld_out >= (int64_t)a->heads * d, "mobi_attention: bad ld_out");
    AttnParams p{};
    p.heads = a->heads;
    p.head_dim = d;
    p.tq = a->tq;
    p.tk = a->tk;
    p.nch = (d + 63) / 64;
    p.dk16 = (d + 15) / 16;
    p.dn = p.dk16 * 16;
    p.ld_out = a->ld_out;
    p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
    p.lse = a->lse;
    MOBI_CHECK(a->lse == nullptr || a->v_rowmajor, "mobi_attention: lse output needs the row-major V kernel (head_dim <= 128)");
    const long long BH = (long long)a->batch * a->heads;
    MOBI_CHECK(BH <= 65535, "mobi_attention: batch*heads=%lld exceeds grid.y", BH);
    if (a->v_rowmajor) {
        MOBI_CHECK(d <= 128, "mobi_attention: the row-major V layout needs head_dim <= 128 (got %d)", d);
        // kernel & 15 == 3 / 4: the third-generation kernel (P staged in shared memory; 4 = its two-tile variant), kept
        // as the A/B arm and for
        // head dims whose O accumulator leaves no TMEM columns for the row sums
        const int sel = a->kernel & 15;
        if (attention4_supports(d) && sel != 3 && sel != 4) return attention4_dispatch(a, p, stream);
        return attention3_dispatch(a, p, stream);
    }
    // shared memory plan: prefer double-buffered K/V and P; fall back to single buffers for wide heads
    auto smem_need = [&](int kvs, int pbs) {
        return (long long)p.nch * ATT_TILE * (1 + kvs) + (long long)kvs * 2 * att_v_tile_bytes(p.dn) +
               (long long)pbs * 2 * ATT_TILE + 256 + 1024;
    };
    const long long limit = 227 * 1024;
    if (smem_need(2, 2) <= limit) {
        p.kv_stages = 2;
        p.p_bufs = 2;
    } else if (smem_need(2, 1) <= limit) {
        p.kv_stages = 2;
        p.p_bufs = 1;
    } else {
        p.kv_stages = 1;
        p.p_bufs = 1;
    }
    const long long smem = smem_need(p.kv_stages, p.p_bufs);
    MOBI_CHECK(smem <= limit, "mobi_attention: head_dim=%d needs %lld bytes of shared memory", d, smem);

    CUtensorMap tmQ, tmK, tmV;
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)a->tq, (uint64_t)BH};
        uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)a->tq * d * 2};
        uint32_t box[3] = {64, ATT_BM, 1};
        if (make_tensor_map_bf16(&tmQ, a->q, 3, dims, strides, box)) return 1;
    }
    {
        uint64_t dims[3] = {(uint64_t)d, (uint64_t)a->tk, (uint64_t)BH};
        uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)a->tk * d * 2};
        uint32_t box[3] = {64, ATT_BKV, 1};
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
        MOBI_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        configured = true;
    }
    dim3 grid((a->tq + ATT_BM - 1) / ATT_BM, (unsigned)BH, 1);
    MOBI_CHECK(BH <= 65535, "mobi_attention: batch*heads=%lld exceeds grid.y", BH);
    MOBI_CUDA(launch_pdl(attention_kernel, grid, dim3(192), (size_t)smem, stream, tmQ, tmK, tmV, p));
    return 0;
}
