// Score-tile kernel of the attention backward (training step): for one batch row and `Z` heads it recomputes, tile by
// tile on tcgen05,
//     S  = q' k^T   (log2 domain, scale * log2(e) folded into q')      and      dP = dO v^T
// into two TMEM accumulators of the same 128 x 128 tile and consumes them in the epilogue WITHOUT ever writing the
// f32 T x T tiles to HBM:
//   STATS pass : per query row the online softmax statistics  rowmax, 1 / rowsum, Delta = sum_j P_ij dP_ij
//   MAIN pass  : P = 2^(S - rowmax) / rowsum ; dS = dscale * P * (dP - Delta) ; writes dS [T, T] row-major and
//                dS^T, P^T [T, T] (bf16): the K-major A operands of dQ = dS k, dK = dS^T q', dV = P^T dO.
// Both passes run the same MMAs on the same operands, so P and Delta are mutually exact (the rows of dS sum to ~1e-7,
// which the cross-modal to_k / to_v weight gradients need; see UNetTrainer.lse_backward for what happens otherwise).
// HBM traffic per head: 3 x T^2 bf16 written (+ read by the three GEMMs) instead of 2 x T^2 f32 written and read twice.
//
// Persistent CTAs own whole (head, 128-query-row) items and walk the key tiles, so the per-row statistics stay in
// registers.  Warp roles as in gemm2.cu: warp 0 TMA producer (Q / dO tiles resident per item, K / V tiles through a
// ring), warp 1 MMA issuer (two TMEM accumulator stages of 256 columns: S | dP), warps 2-9 epilogue (thread = query row,
// two warps per TMEM lane group, each taking 64 of the 128 key columns).
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int AB_BM = 128, AB_BN = 128, AB_BK = 64;
constexpr int AB_TILE = AB_BM * AB_BK * 2;  // 16 KB: one 128 x 64 bf16 tile
constexpr int AB_EPI_WARPS = 16;                       // 4 per TMEM lane group: each takes 32 of the 128 key columns
constexpr int AB_THREADS = 64 + 32 * AB_EPI_WARPS;
constexpr int AB_MAX_NKB = 2;  // head_dim <= 128

struct AttnBwdParams {
    int T, D, nkb, Z, H;   // Z = batch rows x H heads
    float dscale;
    float* stats;          // [Z, T, 3]
    __nv_bfloat16* dS;     // [Z, T, T]
    __nv_bfloat16* dSt;    // [Z, T, T]
    __nv_bfloat16* Pt;     // [Z, T, T]
    int stages;
};

// 2^x on the MUFU pipe without exp2f's range fix-up instructions (inputs are <= 0 here; tiny results flush to zero)
__device__ __forceinline__ float ab_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void ab_bar_sync_epilogue() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <bool STATS>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tiles_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmV,
                      const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nkb = p.nkb;
    const int STAGES = p.stages;
    uint8_t* sQ = smem;                                  // nkb tiles
    uint8_t* sdO = sQ + nkb * AB_TILE;                   // nkb tiles
    uint8_t* sK = sdO + nkb * AB_TILE;                   // STAGES x nkb tiles
    uint8_t* sV = sK + STAGES * nkb * AB_TILE;           // STAGES x nkb tiles
    float* part = reinterpret_cast<float*>(sV + STAGES * nkb * AB_TILE);  // [4][128][3] statistics of the column quarters
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + 4 * 128 * 3);
    uint64_t* full_bar = bars;                // STAGES
    uint64_t* empty_bar = full_bar + 8;       // STAGES
    uint64_t* q_full = empty_bar + 8;         // 1
    uint64_t* q_empty = q_full + 1;           // 1
    uint64_t* tfull_bar = q_empty + 1;        // 2
    uint64_t* tempty_bar = tfull_bar + 2;     // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int m_tiles = p.T / AB_BM, n_tiles = p.T / AB_BN;
    const int items = p.Z * m_tiles;

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
            mbar_init(q_full, 1);
            mbar_init(q_empty, 1);
            for (int s = 0; s < 2; ++s) {
                mbar_init(&tfull_bar[s], 1);
                mbar_init(&tempty_bar[s], AB_EPI_WARPS);
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
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // ---------------- TMA producer
            uint32_t it = 0, li = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
                const int z = item / m_tiles, m_tile = item - z * m_tiles;
                mbar_wait(q_empty, (li & 1) ^ 1);  // the MMAs of the previous item have consumed Q / dO
                mbar_arrive_expect_tx(q_full, 2 * nkb * AB_TILE);
                for (int c = 0; c < nkb; ++c) {
                    tma_load_3d(sQ + c * AB_TILE, &tmQ, q_full, c * AB_BK, m_tile * AB_BM, z);
                    tma_load_4d(sdO + c * AB_TILE, &tmdO, q_full, c * AB_BK, m_tile * AB_BM, z % p.H, z / p.H);
                }
                for (int n = 0; n < n_tiles; ++n, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], 2 * nkb * AB_TILE);
                    for (int c = 0; c < nkb; ++c) {
                        tma_load_3d(sK + (s * nkb + c) * AB_TILE, &tmK, &full_bar[s], c * AB_BK, n * AB_BN, z);
                        tma_load_3d(sV + (s * nkb + c) * AB_TILE, &tmV, &full_bar[s], c * AB_BK, n * AB_BN, z);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ---------------- MMA issuer
            constexpr uint32_t idesc = make_idesc_bf16(AB_BM, AB_BN);
            uint32_t it = 0, lt = 0, li = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++li) {
                mbar_wait(q_full, li & 1);
                tc_fence_after();
                for (int n = 0; n < n_tiles; ++n, ++it, ++lt) {
                    const uint32_t as = lt & 1;
                    mbar_wait(&tempty_bar[as], ((lt >> 1) & 1) ^ 1);
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t dS_t = tmem_base + as * 256, dP_t = dS_t + AB_BN;
                    for (int c = 0; c < nkb; ++c) {
                        const uint64_t qd = make_kmajor_sw128_desc(smem_u32(sQ + c * AB_TILE));
                        const uint64_t kd = make_kmajor_sw128_desc(smem_u32(sK + (s * nkb + c) * AB_TILE));
                        const int ks = (min(AB_BK, p.D - c * AB_BK) + 15) >> 4;  // k-steps that hold head-dim columns
                        for (int k = 0; k < ks; ++k)
                            umma_bf16_ss(dS_t, qd + 2 * k, kd + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                    }
                    for (int c = 0; c < nkb; ++c) {
                        const uint64_t od = make_kmajor_sw128_desc(smem_u32(sdO + c * AB_TILE));
                        const uint64_t vd = make_kmajor_sw128_desc(smem_u32(sV + (s * nkb + c) * AB_TILE));
                        const int ks = (min(AB_BK, p.D - c * AB_BK) + 15) >> 4;
                        for (int k = 0; k < ks; ++k)
                            umma_bf16_ss(dP_t, od + 2 * k, vd + 2 * k, idesc, (c | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                    umma_commit(&tfull_bar[as]);
                }
                umma_commit(q_empty);  // every MMA that reads this item's Q / dO tiles has completed
            }
        }
    } else {
        // ---------------- epilogue warps 2..17: TMEM lane group = warp % 4, column quarter = (warp - 2) / 4
        const int lg = warp & 3;
        const int half = (warp - 2) >> 2;  // 0..3: key columns [32 * half, 32 * half + 32) of the tile
        const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
        const int row_in_tile = lg * 32 + lane;
        uint32_t lt = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int z = item / m_tiles, m_tile = item - z * m_tiles;
            const int row = m_tile * AB_BM + row_in_tile;
            float st_m = 0.f, st_il = 0.f, st_d = 0.f;           // MAIN: statistics of this row
            float m = -INFINITY, l = 0.f, acc = 0.f;             // STATS: running values over this half's columns
            if (!STATS) {
                const float* st = p.stats + ((long long)z * p.T + row) * 3;
                st_m = st[0];
                st_il = st[1];
                st_d = st[2];
            }
            for (int n = 0; n < n_tiles; ++n, ++lt) {
                const uint32_t as = lt & 1;
                mbar_wait(&tfull_bar[as], (lt >> 1) & 1);
                tc_fence_after();
                {
                    const int col0 = half * 32;  // first key column of this warp's chunk inside the tile
                    uint32_t s[32], d[32];
                    tmem_ld32(tmem_base + as * 256 + lane_addr + col0, s);
                    tmem_ld32(tmem_base + as * 256 + AB_BN + lane_addr + col0, d);
                    tmem_ld_wait();
                    if (STATS) {
                        float cm = __uint_as_float(s[0]);
#pragma unroll
                        for (int j = 1; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(s[j]));
                        if (cm > m) {
                            const float r = ab_ex2(m - cm);  // 0 when m == -inf
                            l *= r;
                            acc *= r;
                            m = cm;
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float e = ab_ex2(__uint_as_float(s[j]) - m);
                            l += e;
                            acc = fmaf(e, __uint_as_float(d[j]), acc);
                        }
                    } else {
                        const long long key0 = (long long)n * AB_BN + col0;
                        uint32_t ds2[16], p2[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float p0 = ab_ex2(__uint_as_float(s[j]) - st_m) * st_il;
                            const float p1 = ab_ex2(__uint_as_float(s[j + 1]) - st_m) * st_il;
                            const float g0 = p.dscale * p0 * (__uint_as_float(d[j]) - st_d);
                            const float g1 = p.dscale * p1 * (__uint_as_float(d[j + 1]) - st_d);
                            ds2[j >> 1] = pack_bf16x2(g0, g1);
                            p2[j >> 1] = pack_bf16x2(p0, p1);
                        }
                        // dS row-major: this thread's 32 consecutive key columns = 64 contiguous bytes
                        uint4* drow = reinterpret_cast<uint4*>(p.dS + ((long long)z * p.T + row) * p.T + key0);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            drow[q] = make_uint4(ds2[4 * q], ds2[4 * q + 1], ds2[4 * q + 2], ds2[4 * q + 3]);
                        if (p.dSt == nullptr) {
                            // P row-major (the dK / dV GEMMs read dS and P as MN-major operands: no transposed tiles)
                            uint4* prow = reinterpret_cast<uint4*>(p.Pt + ((long long)z * p.T + row) * p.T + key0);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                prow[q] = make_uint4(p2[4 * q], p2[4 * q + 1], p2[4 * q + 2], p2[4 * q + 3]);
                        } else {
                        // dS^T / P^T: for key column j the 32 lanes hold 32 consecutive query rows; even lanes pair with
                        // their odd neighbour and write one bf16x2 word -> 64 contiguous bytes per (warp, column)
                        const long long tbase = ((long long)z * p.T + key0) * p.T + m_tile * AB_BM + lg * 32 + (lane & ~1);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t wds = ds2[j >> 1], wp = p2[j >> 1];
                            const uint32_t mine_ds = (j & 1) ? (wds >> 16) : (wds & 0xffffu);
                            const uint32_t mine_p = (j & 1) ? (wp >> 16) : (wp & 0xffffu);
                            const uint32_t oth_ds = __shfl_xor_sync(0xffffffffu, mine_ds, 1);
                            const uint32_t oth_p = __shfl_xor_sync(0xffffffffu, mine_p, 1);
                            if ((lane & 1) == 0) {
                                *reinterpret_cast<uint32_t*>(p.dSt + tbase + (long long)j * p.T) = mine_ds | (oth_ds << 16);
                                *reinterpret_cast<uint32_t*>(p.Pt + tbase + (long long)j * p.T) = mine_p | (oth_p << 16);
                            }
                        }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
            }
            if (STATS) {
                float* mine = part + (half * 128 + row_in_tile) * 3;
                mine[0] = m;
                mine[1] = l;
                mine[2] = acc;
                ab_bar_sync_epilogue();
                if (half == 0) {
                    float M = m;
#pragma unroll
                    for (int h = 1; h < 4; ++h) M = fmaxf(M, part[(h * 128 + row_in_tile) * 3]);
                    float L = l * ab_ex2(m - M), A = acc * ab_ex2(m - M);
#pragma unroll
                    for (int h = 1; h < 4; ++h) {
                        const float* oth = part + (h * 128 + row_in_tile) * 3;
                        const float r = ab_ex2(oth[0] - M);
                        L += oth[1] * r;
                        A += oth[2] * r;
                    }
                    float* st = p.stats + ((long long)z * p.T + row) * 3;
                    st[0] = M;
                    st[1] = 1.0f / L;
                    st[2] = A / L;
                }
                ab_bar_sync_epilogue();  // `part` may be overwritten by the next item
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

extern "C" int mobi_attn_bwd_tiles(const mobi_attn_bwd_tiles_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->q && a->k && a->v && a->d_o && a->stats, "mobi_attn_bwd_tiles: null argument");
    MOBI_CHECK(a->stats_only || (a->dS && a->Pt), "mobi_attn_bwd_tiles: the main pass needs dS and Pt (P^T with dSt, row-major P without)");
    MOBI_CHECK(a->tokens > 0 && a->tokens % 128 == 0, "mobi_attn_bwd_tiles: tokens=%d must be a multiple of 128", a->tokens);
    MOBI_CHECK(a->head_dim % 8 == 0 && a->head_dim >= 8 && a->head_dim <= 64 * AB_MAX_NKB,
               "mobi_attn_bwd_tiles: head_dim=%d must be a multiple of 8 in [8, 128]", a->head_dim);
    MOBI_CHECK(a->heads > 0 && a->ld_do % 8 == 0 && a->ld_do >= (int64_t)a->heads * a->head_dim,
               "mobi_attn_bwd_tiles: bad heads / ld_do");
    AttnBwdParams p{};
    p.T = a->tokens;
    p.D = a->head_dim;
    p.nkb = (a->head_dim + 63) / 64;
    p.Z = a->heads * (a->batch_rows > 0 ? a->batch_rows : 1);
    p.H = a->heads;
    p.dscale = a->dscale;
    p.stats = a->stats;
    p.dS = reinterpret_cast<__nv_bfloat16*>(a->dS);
    p.dSt = reinterpret_cast<__nv_bfloat16*>(a->dSt);
    p.Pt = reinterpret_cast<__nv_bfloat16*>(a->Pt);
    const long long fixed = 2ll * p.nkb * AB_TILE + 4 * 128 * 3 * 4 + 512 + 1024;
    const long long per_stage = 2ll * p.nkb * AB_TILE;
    long long stages = (227 * 1024 - fixed) / per_stage;
    if (stages > 6) stages = 6;
    MOBI_CHECK(stages >= 2, "mobi_attn_bwd_tiles: head_dim=%d leaves no room for a K/V ring", a->head_dim);
    p.stages = (int)stages;
    const long long smem = fixed + stages * per_stage;
    CUtensorMap tmQ, tmK, tmV, tmdO;
    const uint64_t T = a->tokens, D = a->head_dim, Z = (uint64_t)p.Z, Hh = a->heads;
    {
        uint64_t dims[3] = {D, T, Z};
        uint64_t strides[2] = {D * 2, T * D * 2};
        uint32_t box[3] = {AB_BK, AB_BM, 1};
        if (make_tensor_map_bf16(&tmQ, a->q, 3, dims, strides, box)) return 1;
        if (make_tensor_map_bf16(&tmK, a->k, 3, dims, strides, box)) return 1;
        if (make_tensor_map_bf16(&tmV, a->v, 3, dims, strides, box)) return 1;
    }
    {
        // token-major dO [batch_rows * T, ld_do]: head h of batch row b = rows [b*T, (b+1)*T), columns [h*D, (h+1)*D)
        uint64_t dims[4] = {D, T, Hh, Z / Hh};
        uint64_t strides[3] = {(uint64_t)a->ld_do * 2, D * 2, T * (uint64_t)a->ld_do * 2};
        uint32_t box[4] = {AB_BK, AB_BM, 1, 1};
        if (make_tensor_map_bf16(&tmdO, a->d_o, 4, dims, strides, box)) return 1;
    }
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(attn_bwd_tiles_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MOBI_CUDA(cudaFuncSetAttribute(attn_bwd_tiles_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    const long long items = (long long)p.Z * (a->tokens / 128);
    const int grid = (int)(items < sm_count() ? items : sm_count());
    attn_bwd_tiles_kernel<true><<<grid, AB_THREADS, smem, stream>>>(tmQ, tmK, tmdO, tmV, p);
    MOBI_CUDA(cudaGetLastError());
    if (!a->stats_only) {
        attn_bwd_tiles_kernel<false><<<grid, AB_THREADS, smem, stream>>>(tmQ, tmK, tmdO, tmV, p);
        MOBI_CUDA(cudaGetLastError());
    }
    return 0;
}
