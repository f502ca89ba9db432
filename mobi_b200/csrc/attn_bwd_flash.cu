// Flash-style attention backward for the training step (CrossAttention.forward under autograd, ldm/modules/attention.py:
// 179-192): dQ', dK and dV of softmax(q' k^T) v WITHOUT writing the T x T score tiles to HBM.  Given the exact row
// statistics (rowmax, 1 / rowsum, Delta) of attn_bwd.cu's statistics pass, two persistent kernels recompute per 128 x 128
// tile, on tcgen05,
//     S = q' k^T  and  dP = dO v^T            (TMEM columns 0-127 and 128-255)
// turn them into P = 2^(S - rowmax) / rowsum and dS = dscale * P * (dP - Delta) in registers (thread = query row), park
// them as bf16 in a 128B-swizzled shared-memory tile, and feed that tile straight back to the tensor core:
//   KEY_OUTER = false : a CTA owns (head, 128 query rows), walks the key tiles,  dQ' += dS . k          (TMEM 256..)
//   KEY_OUTER = true  : a CTA owns (head, 128 key rows),   walks the query tiles, dK += dS^T q', dV += P^T dO  (256.., 384..)
// The staging tile is stored [query row][64-key chunk] exactly like a TMA-loaded K-major tile, so ONE layout serves dS as
// the K-major A operand of dQ and dS^T / P^T as MN-major A operands of dK / dV; the q', k, dO tiles already in shared
// memory for S / dP are re-read as MN-major B operands.  Accumulators leave TMEM once per item as bf16 rows of the
// token-major gradient matrices.  HBM traffic: q', k, v, dO tiles (L2-resident per head) + 3 floats of statistics per row.
//
// Warp roles: warp 0 TMA producer (resident operands per item + a ring for the walked side), warp 1 MMA issuer, warps
// 2-17 epilogue (TMEM lane group = warp % 4 = 32 query rows, column quarter = (warp - 2) / 4 = 32 of the 128 keys).
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int FB_BM = 128, FB_BK = 64;
constexpr int FB_TILE = FB_BM * FB_BK * 2;  // 16 KB: one 128 x 64 bf16 tile
constexpr int FB_STG = 2 * FB_TILE;         // 32 KB: a 128 x 128 bf16 staging tile (two 64-column chunks)
constexpr int FB_EPI_WARPS = 16;
constexpr int FB_THREADS = 64 + 32 * FB_EPI_WARPS;

struct FlashBwdParams {
    int T, D, nkb, Z, H, stages, dpad;
    float dscale;
    const float* stats;     // [Z, T, 3]
    __nv_bfloat16* out0;    // dQ' (KEY_OUTER = false) / dK (true): element (b, t, h, d) at (b * T + t) * ld0 + h * D + d
    __nv_bfloat16* out1;    // dV (KEY_OUTER = true)
    long long ld0, ld1;
};

__device__ __forceinline__ float fb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool KEY_OUTER>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_flash_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmV,
                      const FlashBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nkb = p.nkb;
    const int STAGES = p.stages;
    // resident pair: (q', dO) of the owned query tile, or (k, v) of the owned key tile; ring pair: the walked side
    uint8_t* sR0 = smem;
    uint8_t* sR1 = sR0 + nkb * FB_TILE;
    uint8_t* sW0 = sR1 + nkb * FB_TILE;               // STAGES x nkb tiles
    uint8_t* sW1 = sW0 + STAGES * nkb * FB_TILE;      // STAGES x nkb tiles
    uint8_t* stg_dS = sW1 + STAGES * nkb * FB_TILE;   // [128 query rows][128 keys] bf16, 128B swizzle
    uint8_t* stg_P = stg_dS + FB_STG;                 // KEY_OUTER only
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_dS + (KEY_OUTER ? 2 : 1) * FB_STG);
    uint64_t* full_bar = bars;              // STAGES (<= 8)
    uint64_t* empty_bar = full_bar + 8;     // STAGES
    uint64_t* r_full = empty_bar + 8;
    uint64_t* r_empty = r_full + 1;
    uint64_t* sdp_full = r_empty + 1;       // MMA -> epilogue: S and dP of this tile are in TMEM
    uint64_t* sdp_empty = sdp_full + 1;     // epilogue -> MMA: S / dP columns have been read
    uint64_t* stg_full = sdp_empty + 1;     // epilogue -> MMA: dS (and P) tiles are staged in shared memory
    uint64_t* stg_empty = stg_full + 1;     // MMA -> epilogue: the accumulate MMAs have consumed the staged tiles
    uint64_t* acc_full = stg_empty + 1;     // MMA -> epilogue: the accumulators of this item are final
    uint64_t* acc_empty = acc_full + 1;     // epilogue -> MMA: accumulators drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int tiles = p.T / FB_BM;          // query tiles == key tiles
    const int items = p.Z * tiles;

    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmdO);
            tma_prefetch_desc(&tmV);
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_init(r_full, 1);
            mbar_init(r_empty, 1);
            mbar_init(sdp_full, 1);
            mbar_init(sdp_empty, FB_EPI_WARPS);
            mbar_init(stg_full, FB_EPI_WARPS);
            mbar_init(stg_empty, 1);
            mbar_init(acc_full, 1);
            mbar_init(acc_empty, FB_EPI_WARPS);
            fence_barrier_init();
        }
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ---------------- TMA producer
            uint32_t it = 0, li = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
                const int z = item / tiles, own = item - z * tiles;
                mbar_wait(r_empty, (li & 1) ^ 1);  // every MMA of the previous item has read its resident tiles
                mbar_arrive_expect_tx(r_full, 2 * nkb * FB_TILE);
                for (int c = 0; c < nkb; ++c) {
                    if (KEY_OUTER) {
                        tma_load_3d(sR0 + c * FB_TILE, &tmK, r_full, c * FB_BK, own * FB_BM, z);
                        tma_load_3d(sR1 + c * FB_TILE, &tmV, r_full, c * FB_BK, own * FB_BM, z);
                    } else {
                        tma_load_3d(sR0 + c * FB_TILE, &tmQ, r_full, c * FB_BK, own * FB_BM, z);
                        tma_load_4d(sR1 + c * FB_TILE, &tmdO, r_full, c * FB_BK, own * FB_BM, z % p.H, z / p.H);
                    }
                }
                for (int i = 0; i < tiles; ++i, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], 2 * nkb * FB_TILE);
                    for (int c = 0; c < nkb; ++c) {
                        if (KEY_OUTER) {
                            tma_load_3d(sW0 + (s * nkb + c) * FB_TILE, &tmQ, &full_bar[s], c * FB_BK, i * FB_BM, z);
                            tma_load_4d(sW1 + (s * nkb + c) * FB_TILE, &tmdO, &full_bar[s], c * FB_BK, i * FB_BM, z % p.H, z / p.H);
                        } else {
                            tma_load_3d(sW0 + (s * nkb + c) * FB_TILE, &tmK, &full_bar[s], c * FB_BK, i * FB_BM, z);
                            tma_load_3d(sW1 + (s * nkb + c) * FB_TILE, &tmV, &full_bar[s], c * FB_BK, i * FB_BM, z);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ---------------- MMA issuer
            constexpr uint32_t idesc_s = make_idesc_bf16(FB_BM, 128);
            // accumulate products: N = padded head dim; B always MN-major, A MN-major when it is a transposed tile
            const uint32_t idesc_acc = make_idesc_bf16(FB_BM, (p.D + 15) & ~15) | (1u << 16) | (KEY_OUTER ? 1u << 15 : 0u);
            const uint32_t S_t = tmem_base, dP_t = tmem_base + 128, acc0_t = tmem_base + 256, acc1_t = tmem_base + 384;
            // S = q' k^T and dP = dO v^T of walked tile `t` (global tile counter) into TMEM columns 0-255
            auto issue_scores = [&](uint32_t t) {
                const int s = t % STAGES;
                mbar_wait(&full_bar[s], (t / STAGES) & 1);
                mbar_wait(sdp_empty, (t & 1) ^ 1);  // the epilogue holds the previous tile's S / dP in registers
                tc_fence_after();
                const uint8_t* w0 = sW0 + s * nkb * FB_TILE;
                const uint8_t* w1 = sW1 + s * nkb * FB_TILE;
                const uint8_t* q_t = KEY_OUTER ? w0 : sR0;
                const uint8_t* k_t = KEY_OUTER ? sR0 : w0;
                const uint8_t* do_t = KEY_OUTER ? w1 : sR1;
                const uint8_t* v_t = KEY_OUTER ? sR1 : w1;
                // only the 16-column k-steps that hold head-dim columns (3 of 4 at D = 40; the rest is TMA zero fill)
                for (int c = 0; c < nkb; ++c) {
                    const uint64_t qd = make_kmajor_sw128_desc(smem_u32(q_t + c * FB_TILE));
                    const uint64_t kd = make_kmajor_sw128_desc(smem_u32(k_t + c * FB_TILE));
                    const int ks = (min(FB_BK, p.D - c * FB_BK) + 15) >> 4;
                    for (int k = 0; k < ks; ++k) umma_bf16_ss(S_t, qd + 2 * k, kd + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
                }
                for (int c = 0; c < nkb; ++c) {
                    const uint64_t od = make_kmajor_sw128_desc(smem_u32(do_t + c * FB_TILE));
                    const uint64_t vd = make_kmajor_sw128_desc(smem_u32(v_t + c * FB_TILE));
                    const int ks = (min(FB_BK, p.D - c * FB_BK) + 15) >> 4;
                    for (int k = 0; k < ks; ++k) umma_bf16_ss(dP_t, od + 2 * k, vd + 2 * k, idesc_s, (c | k) != 0 ? 1u : 0u);
                }
                umma_commit(sdp_full);
            };
            // With a ring of >= 2 stages the scores of tile i + 1 are issued BEFORE the accumulate products of tile i: the
            // S / dP columns are free as soon as the epilogue has loaded them into registers, so the tensor core computes
            // the next scores while the epilogue still exponentiates / packs / stages the current tile.  (With one ring
            // stage tile i + 1 cannot be loaded before the accumulate products of tile i release the stage.)
            const bool ahead = STAGES >= 2;
            uint32_t it = 0, li = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
                mbar_wait(r_full, li & 1);
                mbar_wait(acc_empty, (li & 1) ^ 1);  // the epilogue has drained the previous item's accumulators
                tc_fence_after();
                if (ahead) issue_scores(it);
                for (int i = 0; i < tiles; ++i, ++it) {
                    const int s = it % STAGES;
                    if (!ahead) issue_scores(it);
                    else if (i + 1 < tiles) issue_scores(it + 1);
                    const uint8_t* w0 = sW0 + s * nkb * FB_TILE;
                    const uint8_t* w1 = sW1 + s * nkb * FB_TILE;
                    // the epilogue turns S / dP into bf16 dS (and P) tiles in shared memory
                    mbar_wait(stg_full, it & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {  // 128 contraction rows (keys or query rows), 16 per MMA
                        const uint32_t accf = (i | k) != 0 ? 1u : 0u;
                        if (KEY_OUTER) {
                            // dK[keys, d] += dS^T q' ; dV[keys, d] += P^T dO : A = staged tile read MN-major (M = keys: two
                            // 64-key chunks 16 KB apart), B = walked q' / dO tile read MN-major (N = d chunks one tile apart)
                            const uint64_t a_ds = make_mnmajor_sw128_desc(smem_u32(stg_dS) + k * 2048, FB_TILE);
                            const uint64_t a_p = make_mnmajor_sw128_desc(smem_u32(stg_P) + k * 2048, FB_TILE);
                            const uint64_t b_q = make_mnmajor_sw128_desc(smem_u32(w0) + k * 2048, FB_TILE);
                            const uint64_t b_do = make_mnmajor_sw128_desc(smem_u32(w1) + k * 2048, FB_TILE);
                            umma_bf16_ss(acc0_t, a_ds, b_q, idesc_acc, accf);
                            umma_bf16_ss(acc1_t, a_p, b_do, idesc_acc, accf);
                        } else {
                            // dQ'[q, d] += dS k : A = staged dS read K-major (16 keys = 32 B inside a 64-key chunk),
                            // B = walked k tile read MN-major
                            const uint64_t a_ds = make_kmajor_sw128_desc(smem_u32(stg_dS + (k >> 2) * FB_TILE)) + 2 * (k & 3);
                            const uint64_t b_k = make_mnmajor_sw128_desc(smem_u32(w0) + k * 2048, FB_TILE);
                            umma_bf16_ss(acc0_t, a_ds, b_k, idesc_acc, accf);
                        }
                    }
                    umma_commit(stg_empty);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(acc_full);
                umma_commit(r_empty);
            }
        }
    } else {
        // ---------------- epilogue warps 2..17
        const int lg = warp & 3;
        const int quarter = (warp - 2) >> 2;  // key columns [32 * quarter, +32) of the tile
        const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
        const int row_in_tile = lg * 32 + lane;
        const int col0 = quarter * 32;
        // staging address of this thread's 64 bytes: chunk = col0 / 64, 16-byte units (col0 % 64) / 8 .. + 3, XOR (row & 7)
        const uint32_t stg_off = (col0 >> 6) * FB_TILE + row_in_tile * 128;
        const int unit0 = (col0 & 63) >> 3;
        uint32_t it = 0, li = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
            const int z = item / tiles, own = item - z * tiles;
            float st_m = 0.f, st_il = 0.f, st_d = 0.f;
            if (!KEY_OUTER) {
                const float* st = p.stats + ((long long)z * p.T + own * FB_BM + row_in_tile) * 3;
                st_m = st[0];
                st_il = st[1];
                st_d = st[2];
            }
            for (int i = 0; i < tiles; ++i, ++it) {
                if (KEY_OUTER) {  // the query rows change with every walked tile
                    const float* st = p.stats + ((long long)z * p.T + i * FB_BM + row_in_tile) * 3;
                    st_m = st[0];
                    st_il = st[1];
                    st_d = st[2];
                }
                mbar_wait(sdp_full, it & 1);
                tc_fence_after();
                uint32_t s[32], d[32];
                tmem_ld32(tmem_base + lane_addr + col0, s);
                tmem_ld32(tmem_base + 128 + lane_addr + col0, d);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sdp_empty);
                // P = 2^(S - m) / l = 2^(S - (m + log2 l)); dS = P * (dscale * dP - dscale * Delta): one add, one MUFU, one
                // FMA and one multiply per element
                const float m_l = st_m - __log2f(st_il), neg_sd = -p.dscale * st_d;
                uint32_t ds2[16], p2[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = fb_ex2(__uint_as_float(s[j]) - m_l);
                    const float p1 = fb_ex2(__uint_as_float(s[j + 1]) - m_l);
                    const float g0 = p0 * fmaf(p.dscale, __uint_as_float(d[j]), neg_sd);
                    const float g1 = p1 * fmaf(p.dscale, __uint_as_float(d[j + 1]), neg_sd);
                    ds2[j >> 1] = pack_bf16x2(g0, g1);
                    p2[j >> 1] = pack_bf16x2(p0, p1);
                }
                mbar_wait(stg_empty, (it & 1) ^ 1);  // the accumulate MMAs of the previous tile have read the staging tiles
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t off = stg_off + (((unit0 + u) ^ (row_in_tile & 7)) << 4);
                    *reinterpret_cast<uint4*>(stg_dS + off) = make_uint4(ds2[4 * u], ds2[4 * u + 1], ds2[4 * u + 2], ds2[4 * u + 3]);
                    if (KEY_OUTER)
                        *reinterpret_cast<uint4*>(stg_P + off) = make_uint4(p2[4 * u], p2[4 * u + 1], p2[4 * u + 2], p2[4 * u + 3]);
                }
                fence_proxy_async();  // generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(stg_full);
            }
            // ---- accumulators of this item -> bf16 rows of the token-major gradient matrices
            mbar_wait(acc_full, li & 1);
            tc_fence_after();
            if (col0 < p.dpad) {
                const int b = z / p.H, h = z - b * p.H;
                const long long trow = (long long)b * p.T + own * FB_BM + row_in_tile;
#pragma unroll
                for (int a = 0; a < (KEY_OUTER ? 2 : 1); ++a) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + 256 + a * 128 + lane_addr + col0, v);
                    tmem_ld_wait();
                    __nv_bfloat16* dst = (a == 0 ? p.out0 + trow * p.ld0 : p.out1 + trow * p.ld1) + h * p.D + col0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (col0 + 8 * u < p.D) {
                            uint4 w;
                            w.x = pack_bf16x2(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1]));
                            w.y = pack_bf16x2(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3]));
                            w.z = pack_bf16x2(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5]));
                            w.w = pack_bf16x2(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7]));
                            reinterpret_cast<uint4*>(dst)[u] = w;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
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

extern "C" int mobi_attn_bwd_flash(const mobi_attn_bwd_flash_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->q && a->k && a->v && a->d_o && a->stats && a->dq && a->dk && a->dv, "mobi_attn_bwd_flash: null argument");
    MOBI_CHECK(a->tokens > 0 && a->tokens % 128 == 0, "mobi_attn_bwd_flash: tokens=%d must be a multiple of 128", a->tokens);
    MOBI_CHECK(a->head_dim % 8 == 0 && a->head_dim >= 8 && a->head_dim <= 128,
               "mobi_attn_bwd_flash: head_dim=%d must be a multiple of 8 in [8, 128]", a->head_dim);
    const int64_t inner = (int64_t)a->heads * a->head_dim;
    MOBI_CHECK(a->heads > 0 && a->ld_do % 8 == 0 && a->ld_do >= inner && a->ld_dq % 8 == 0 && a->ld_dq >= inner &&
                   a->ld_dk % 8 == 0 && a->ld_dk >= inner && a->ld_dv % 8 == 0 && a->ld_dv >= inner,
               "mobi_attn_bwd_flash: row strides must be multiples of 8 and >= heads * head_dim");
    MOBI_CHECK(reinterpret_cast<uintptr_t>(a->dq) % 16 == 0 && reinterpret_cast<uintptr_t>(a->dk) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(a->dv) % 16 == 0,
               "mobi_attn_bwd_flash: outputs must be 16-byte aligned");
    FlashBwdParams p{};
    p.T = a->tokens;
    p.D = a->head_dim;
    p.nkb = (a->head_dim + 63) / 64;
    p.dpad = 64 * p.nkb;
    p.Z = a->heads * (a->batch_rows > 0 ? a->batch_rows : 1);
    p.H = a->heads;
    p.dscale = a->dscale;
    p.stats = a->stats;
    CUtensorMap tmQ, tmK, tmV, tmdO;
    const uint64_t T = a->tokens, D = a->head_dim, Z = (uint64_t)p.Z, Hh = a->heads;
    {
        uint64_t dims[3] = {D, T, Z};
        uint64_t strides[2] = {D * 2, T * D * 2};
        uint32_t box[3] = {FB_BK, FB_BM, 1};
        if (make_tensor_map_bf16(&tmQ, a->q, 3, dims, strides, box)) return 1;
        if (make_tensor_map_bf16(&tmK, a->k, 3, dims, strides, box)) return 1;
        if (make_tensor_map_bf16(&tmV, a->v, 3, dims, strides, box)) return 1;
    }
    {
        uint64_t dims[4] = {D, T, Hh, Z / Hh};
        uint64_t strides[3] = {(uint64_t)a->ld_do * 2, D * 2, T * (uint64_t)a->ld_do * 2};
        uint32_t box[4] = {FB_BK, FB_BM, 1, 1};
        if (make_tensor_map_bf16(&tmdO, a->d_o, 4, dims, strides, box)) return 1;
    }
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(attn_bwd_flash_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MOBI_CUDA(cudaFuncSetAttribute(attn_bwd_flash_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    const long long items = (long long)p.Z * (a->tokens / 128);
    const int grid = (int)(items < sm_count() ? items : sm_count());
    const long long per_stage = 2ll * p.nkb * FB_TILE;
    for (int pass = 0; pass < 2; ++pass) {
        const bool key_outer = pass == 1;
        const long long fixed = 2ll * p.nkb * FB_TILE + (key_outer ? 2 : 1) * FB_STG + 512 + 1024;
        long long stages = (227 * 1024 - fixed) / per_stage;
        if (stages > 4) stages = 4;
        MOBI_CHECK(stages >= 1, "mobi_attn_bwd_flash: head_dim=%d leaves no room for the operand ring", a->head_dim);
        p.stages = (int)stages;
        const long long smem = fixed + stages * per_stage;
        if (key_outer) {
            p.out0 = reinterpret_cast<__nv_bfloat16*>(a->dk);
            p.out1 = reinterpret_cast<__nv_bfloat16*>(a->dv);
            p.ld0 = a->ld_dk;
            p.ld1 = a->ld_dv;
            attn_bwd_flash_kernel<true><<<grid, FB_THREADS, smem, stream>>>(tmQ, tmK, tmdO, tmV, p);
        } else {
            p.out0 = reinterpret_cast<__nv_bfloat16*>(a->dq);
            p.out1 = nullptr;
            p.ld0 = a->ld_dq;
            p.ld1 = 0;
            attn_bwd_flash_kernel<false><<<grid, FB_THREADS, smem, stream>>>(tmQ, tmK, tmdO, tmV, p);
        }
        MOBI_CUDA(cudaGetLastError());
    }
    return 0;
}
