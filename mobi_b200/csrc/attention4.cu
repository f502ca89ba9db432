// Fused attention, fourth generation (head_dim <= 128, V in its natural [keys, d] layout): what the UNet runs.
//
// ncu on the third generation at the level-0 shape (d = 40, T = 4096; profiles/r01/ncu_full_gemm2_attention3.md and the
// raw counters of that capture) showed three co-limiters at ~65 % each: the shared-memory data pipe (957 wavefronts per
// 128 x 128 score tile: 640 tensor-core operand reads + 317 P stores), MUFU (768 clk) and issue slots (985 clk) for a
// tile that took 1500 clk.  This kernel removes most of the shared-memory traffic and a part of the instructions:
//   * P never touches shared memory: the softmax threads write their bf16 probabilities back into the TMEM columns the
//     scores came from (tcgen05.st) and the PV product reads its A operand from TMEM (tcgen05.mma with [a_tmem]):
//     no P stores, no P operand reads, no generic->async proxy fence, no P buffers (32 KB of shared memory freed);
//   * the row sums come from the tensor core: a second, N = 16 MMA per k-step multiplies the same P by a tile of ones
//     into 16 columns next to the O accumulator, so l is the fp32 sum of exactly the bf16 probabilities the PV product
//     used, it is rescaled together with O, and the softmax loop loses its packed add per pair;
//   * the role loops are lean: descriptors are built once and advanced by adds on their address field, mbarriers are
//     addressed through precomputed shared-window addresses, the head dim is a template parameter (full unrolling).
// Two TMEM plans (template DEC):
//   DEC = 1 (what runs): P_t has its own 32 columns, so S_t(j + 1) = Q K(j + 1)^T is issued as soon as the softmax
//     threads have pulled S_t(j) into registers and runs UNDER the exponentials of block j; 160 columns per tile leave
//     room for NT = 3 tiles of 128 query rows per CTA (head_dim <= 48) or 2 (head_dim <= 112).  A first version (DEC = 0,
//     NT = 4, P written over the S columns, S(j + 1) issued after PV(j) by the same thread) left the softmax warps
//     waiting for the MMA round trip 36 % of their time (ncu source page, profiles/r02/): 1.25 ms vs 1.46 ms for
//     attention3 at the level-0 shape, but still latency-bound.
// Everything else is the attention3 design: NT query tiles of 128 rows per CTA (thread = query row = TMEM lane), all
// tiles share one K/V ring, every tile has its own MMA-issuer warp, V is an MN-major B operand, stale running maximum,
// a share of the exponentials on the FMA pipe.
// TMEM: S_t at columns t*BKV, P_t over it (DEC = 0) or in its own 32 columns behind the S tiles, O_t at the next multiple
// of O_STRIDE with l_t in the 16 columns after O_t's DN columns (NT = 3: S 0..191, P 192..287, O 320..511).
#include "../../include/mobi_b200.h"
#include "attention_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int A4_CHUNK = 128 * 128;  // bytes of a 128-row x 64-col bf16 chunk
constexpr int A4_ONES_BYTES = 2048;  // 16 rows x 64 keys of bf16 1.0 (K-major B operand of the row-sum MMA)

__device__ __forceinline__ float a4_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for a pair on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, relative error 1e-4; x <= 8)
__device__ __forceinline__ void a4_ex2_poly2(uint64_t x, float& e0, float& e1) {
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

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 16 bf16 = 8 columns of packed pairs) is read from TMEM.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// mbarrier operations on precomputed shared-window addresses: the role loops below run on the critical path of every
// key block, so nothing in them converts a generic pointer (S2UR / cvta sequences) or rebuilds a descriptor from scratch
__device__ __forceinline__ bool a4_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void a4_wait(uint32_t bar, uint32_t parity) {
    if (a4_try_wait(bar, parity)) return;
    long long start = 0;
#pragma unroll 1
    for (uint32_t spins = 1;; ++spins) {
        if (a4_try_wait(bar, parity)) return;
        if ((spins & 255u) == 0u) {  // bounded: a protocol bug traps instead of hanging the GPU
            const long long now = clock64();
            if (start == 0) start = now;
            else if (now - start > MOBI_WAIT_LIMIT_CYCLES) asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void a4_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void a4_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// SPLIT (tuning arm, kernel & 15 == 6; NOT what runs): TWO softmax warps per 32-row TMEM lane group (8 per tile), each
// turns ITS half of the 64 keys of a block into probabilities; the row maximum is agreed on by swapping partial maxima
// through shared memory and a 64-thread named barrier per block.  The idea: warps issue in order, so with 3 of them per SM
// sub-partition the MUFU pipe (61 % busy, ncu) and the FMA pipe wait on each other, and 6 should interleave better.
// Measured at the level-0 shape (32 rows): 1.34-1.37 ms against 1.22 ms unsplit (1.51 ms when both warps re-read the
// whole score row instead of exchanging the maximum); head_dim 80: 0.151 vs 0.153 ms.  More warps do not pay: the
// per-block fixed costs (waits, TMEM round trips, the exchange) double while the exponentials per warp halve.
template <int NT, int SPLIT = 0>
struct A4Cfg {
    static constexpr int ROLE_WARPS = NT <= 3 ? 4 : 8;
    static constexpr int SW = SPLIT ? 8 : 4;             // softmax warps per query tile
    static constexpr int WARPS = SW * NT + ROLE_WARPS;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int SOFTMAX_REGS = SPLIT ? 0 : (NT == 4 ? 96 : (NT == 3 ? 152 : 208));  // 0: no setmaxnreg
    static constexpr int O_STRIDE = NT >= 3 ? 64 : 128;  // TMEM columns per O accumulator incl. the 16 row-sum columns
};

constexpr uint32_t A4_ROLE_SLEEP_NS = 64;

template <int NT, int BKV, int KVS, int POLY, int DK16, int DEC, int SPLIT>
__global__ void __launch_bounds__(A4Cfg<NT, SPLIT>::THREADS, 1)
attention4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using Cfg = A4Cfg<NT, SPLIT>;
    static_assert(BKV == 64, "one 64-key block = 32 TMEM columns of packed probabilities");
    static_assert(!SPLIT || DEC, "the split softmax is built on the decoupled TMEM plan");
    constexpr int DN = DK16 * 16;                      // head_dim rounded up to the MMA K / N granularity
    constexpr int NCH = (DN + 63) / 64;                // 64-wide chunks of the head dim
    constexpr int K_CHUNK = BKV * 128;                 // bytes of a BKV-row x 64-col bf16 chunk of K or V
    constexpr int O_STRIDE = Cfg::O_STRIDE;
    // TMEM columns: S_t (fp32) at t*BKV; P_t (packed bf16 pairs) over the first half of S_t, or, when decoupled, in its
    // own BKV/2 columns behind all S tiles; the O accumulators start at the next multiple of their stride (accumulator
    // bases are kept aligned to the accumulator width)
    constexpr int P_BASE = DEC ? NT * BKV : 0;
    constexpr int P_STRIDE = DEC ? BKV / 2 : BKV;
    constexpr int O_BASE = ((NT * BKV + (DEC ? NT * BKV / 2 : 0)) + O_STRIDE - 1) / O_STRIDE * O_STRIDE;
    constexpr int BASE = Cfg::SW * NT;                 // first role warp
    constexpr int ROLE_REGS = NT == 4 ? 40 : 56;
    static_assert(O_BASE + NT * O_STRIDE <= 512 && DN + 16 <= O_STRIDE, "TMEM budget");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int q_tile_bytes = NCH * A4_CHUNK;
    constexpr int kv_bytes = NCH * K_CHUNK;            // K and V blocks have the same footprint
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + NT * q_tile_bytes;
    uint8_t* sV = sK + KVS * kv_bytes;
    uint8_t* sOnes = sV + KVS * kv_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + A4_ONES_BYTES);
    uint64_t* q_full = bars;
    uint64_t* s_full = bars + 1;
    uint64_t* p_full = s_full + NT;
    uint64_t* o_done = p_full + NT;   // every PV(j) of the tile commits here (DEC: also "P buffer free"; else the last one)
    uint64_t* s_free = o_done + NT;   // DEC only: the softmax threads have S_t in registers
    uint64_t* kv_full = s_free + NT;
    uint64_t* kv_empty = kv_full + KVS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + KVS);
    float* sXch = reinterpret_cast<float*>(bars + 64);   // SPLIT: per (tile, lane group) 2 x 2 x 32 partial maxima

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (128 * NT);
    const int bh = blockIdx.y;
    const int nblk = (p.tk + BKV - 1) / BKV;
    const int ntiles = min(NT, (p.tq - q0 + 127) / 128);

    // the tile of ones (any layout of an all-ones tile is the same tile)
    for (int i = threadIdx.x; i < A4_ONES_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3f803f80u;
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
                mbar_init(&p_full[i], 32 * Cfg::SW);
                mbar_init(&o_done[i], 1);
                mbar_init(&s_free[i], 32 * Cfg::SW);
            }
            fence_barrier_init();
        }
    } else if (warp == BASE + 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async();  // the ones tile (generic-proxy stores) is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();  // barriers, TMEM and the ones tile were set up under the tail of the previous kernel (q / k / v producer)
    pdl_trigger();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= BASE) {
        // donate registers to the softmax warpgroups
        if constexpr (SPLIT) {
        } else if constexpr (ROLE_REGS == 40) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == BASE) {
            if (elect_one()) {
                // ---------------- TMA producer
                mbar_arrive_expect_tx(q_full, ntiles * q_tile_bytes);
                for (int t = 0; t < ntiles; ++t)
                    for (int c = 0; c < NCH; ++c)
                        tma_load_3d(sQ + t * q_tile_bytes + c * A4_CHUNK, &tmQ, q_full, c * 64, q0 + t * 128, bh);
                for (int j = 0; j < nblk; ++j) {
                    const int s = j % KVS;
                    mbar_wait_backoff(&kv_empty[s], ((j / KVS) & 1) ^ 1, A4_ROLE_SLEEP_NS);
                    mbar_arrive_expect_tx(&kv_full[s], 2 * kv_bytes);
                    for (int c = 0; c < NCH; ++c) {
                        tma_load_3d(sK + s * kv_bytes + c * K_CHUNK, &tmK, &kv_full[s], c * 64, j * BKV, bh);
                        tma_load_3d(sV + s * kv_bytes + c * K_CHUNK, &tmV, &kv_full[s], c * 64, j * BKV, bh);
                    }
                }
            }
        } else if (warp - (BASE + 1) < ntiles) {
            if (elect_one()) {
                // ---------------- MMA issuer of tile t.  All MMAs of the tile come from this thread and complete in issue
                // order: S(j + 1) may overwrite the columns PV(j) read its probabilities from, and the commit that follows
                // S(j + 1) also covers PV(j), which is what lets the softmax threads rescale O after it.
                const int t = warp - (BASE + 1);
                constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV);
                constexpr uint32_t idesc_o = make_idesc_bf16(128, DN) | (1u << 16);  // B operand (V) is MN-major
                constexpr uint32_t idesc_l = make_idesc_bf16(128, 16);               // B operand: the ones tile, K-major
                const uint32_t s_tmem = tmem_base + t * BKV;
                const uint32_t o_tmem = tmem_base + O_BASE + t * O_STRIDE;
                const uint32_t l_tmem = o_tmem + DN;
                // descriptors of stage 0 / chunk 0 / k-step 0: everything else is an add on the 14-bit start-address field
                // (16-byte units; all operands live below 228 KB, so the field never carries)
                const uint64_t ones_desc = make_kmajor_sw128_desc(smem_u32(sOnes));
                const uint64_t q_desc = make_kmajor_sw128_desc(smem_u32(sQ + t * q_tile_bytes));
                const uint64_t k_desc = make_kmajor_sw128_desc(smem_u32(sK));
                const uint64_t v_desc = make_mnmajor_sw128_desc(smem_u32(sV), K_CHUNK);
                const uint32_t b_s_full = smem_u32(&s_full[t]), b_p_full = smem_u32(&p_full[t]);
                const uint32_t b_kv_full = smem_u32(kv_full), b_kv_empty = smem_u32(kv_empty);
                auto issue_S = [&](uint32_t stage_off) {  // S_t = Q_t K(stage)^T
#pragma unroll
                    for (int k = 0; k < DK16; ++k) {
                        const uint32_t off = (k >> 2) * (A4_CHUNK >> 4) + 2 * (k & 3);
                        const uint32_t koff = (k >> 2) * (K_CHUNK >> 4) + 2 * (k & 3);
                        umma_bf16_ss(s_tmem, q_desc + off, k_desc + stage_off + koff, idesc_s, k != 0 ? 1u : 0u);
                    }
                    a4_commit(b_s_full);
                };
                mbar_wait_backoff(q_full, 0, A4_ROLE_SLEEP_NS);
                a4_wait(b_kv_full, 0);
                tc_fence_after();
                issue_S(0);
                const uint32_t b_o_done = smem_u32(&o_done[t]), b_s_free = smem_u32(&s_free[t]);
                const uint32_t p_tmem = tmem_base + P_BASE + t * P_STRIDE;
                uint32_t s = 0, kv_par = 0;
                for (int j = 0; j < nblk; ++j) {
                    uint32_t s1 = s + 1, kv_par1 = kv_par;
                    if (s1 == KVS) {
                        s1 = 0;
                        kv_par1 ^= 1;
                    }
                    if constexpr (DEC) {
                        if (j + 1 < nblk) {  // S(j + 1) under the exponentials of block j
                            a4_wait(b_kv_full + s1 * 8, kv_par1);
                            a4_wait(b_s_free, j & 1);
                            tc_fence_after();
                            issue_S(s1 * (kv_bytes >> 4));
                        }
                    }
                    a4_wait(b_p_full, j & 1);
                    tc_fence_after();
                    const uint64_t vd = v_desc + s * (kv_bytes >> 4);
                    const uint32_t acc = j != 0 ? 1u : 0u;
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k) {  // 16 keys per MMA: A = 8 TMEM columns of packed pairs
                        umma_bf16_ts(o_tmem, p_tmem + k * 8, vd + k * (2048 >> 4), idesc_o, k != 0 ? 1u : acc);
                        umma_bf16_ts(l_tmem, p_tmem + k * 8, ones_desc, idesc_l, k != 0 ? 1u : acc);
                    }
                    if constexpr (DEC) a4_commit(b_o_done);
                    a4_commit(b_kv_empty + s * 8);  // this tile is done with K(j) / V(j)
                    if constexpr (!DEC) {
                        if (j + 1 < nblk) {
                            a4_wait(b_kv_full + s1 * 8, kv_par1);
                            tc_fence_after();
                            issue_S(s1 * (kv_bytes >> 4));
                        } else {
                            a4_commit(b_o_done);
                        }
                    }
                    s = s1;
                    kv_par = kv_par1;
                }
            }
        }
    } else if constexpr (SPLIT) {
        // ---------------- split softmax: warps [8t, 8t + 8) serve tile t; warp & 3 = TMEM lane group, (warp >> 2) & 1 =
        // which 32 keys of the block this warp exponentiates
        const int t = warp >> 3;
        if (t < ntiles) {
            const int lg = warp & 3, hf = (warp >> 2) & 1;
            const int row = lg * 32 + lane;
            const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
            const uint32_t tS = tmem_base + t * BKV + lane_addr;
            const uint32_t tP = tmem_base + P_BASE + t * P_STRIDE + lane_addr + hf * 16;
            float* xch = sXch + (t * 4 + lg) * 128;        // [parity][half][lane]: double-buffered over j
            const int pair_bar = 1 + t * 4 + lg;           // named barrier of this (tile, lane group) pair (0 = __syncthreads)
            const uint32_t tO = tmem_base + O_BASE + t * O_STRIDE + lane_addr;
            const uint32_t b_s_full = smem_u32(&s_full[t]), b_p_full = smem_u32(&p_full[t]);
            const uint32_t b_o_done = smem_u32(&o_done[t]), b_s_free = smem_u32(&s_free[t]);
            float m_used = -INFINITY;
            for (int j = 0; j < nblk; ++j) {
                uint32_t sr[32];
                a4_wait(b_s_full, j & 1);
                tc_fence_after();
                tmem_ld32(tS + hf * 32, sr);                 // this warp's 32 keys
                tmem_ld_wait();
                const int valid = p.tk - j * BKV;  // keys of this block that exist (>= BKV except in the last block)
                tc_fence_before();
                a4_arrive(b_s_free);  // the S columns may be overwritten by the next QK^T
                if (valid < BKV) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (hf * 32 + i >= valid) sr[i] = 0xff800000u;  // -inf
                }
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(sr[i + 2 * u]), __uint_as_float(sr[i + 2 * u + 1])));
                }
                const float mx_own = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                // the row maximum of the whole block: swap partial maxima with the partner warp (same rows, other keys)
                // through shared memory and a 64-thread named barrier; both warps then hold the identical value
                xch[(j & 1) * 64 + hf * 32 + lane] = mx_own;
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                const float mx = fmaxf(mx_own, xch[(j & 1) * 64 + (hf ^ 1) * 32 + lane]);
                if (j == 0) {
                    m_used = mx;
                } else {
                    const bool need = mx > m_used + 8.0f;
                    if (__any_sync(0xffffffffu, need)) {
                        // both warps of the lane group take the same decision from the same scores; each rescales its
                        // share of the O | l chunks once PV(j - 1) is complete
                        a4_wait(b_o_done, (j & 1) ^ 1);
                        tc_fence_after();
                        const float m_new = fmaxf(m_used, mx);
                        const float alpha = a4_ex2(m_used - m_new);
                        m_used = m_new;
#pragma unroll 1
                        for (int c = hf * 16; c < DN + 16; c += 32) {
                            uint32_t r[16];
                            tmem_ld16(tO + c, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            tmem_st16(tO + c, r);
                        }
                    }
                }
                const uint64_t m2 = pack2(m_used, m_used);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const uint64_t x = sub2(pack2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), m2);
                    float e0, e1;
                    constexpr int kPolySlots[9] = {0x00, 0x08, 0x22, 0x2a, 0xaa, 0xab, 0xbb, 0xbf, 0xff};
                    if ((kPolySlots[POLY] >> ((c >> 1) & 7)) & 1) {
                        a4_ex2_poly2(x, e0, e1);
                    } else {
                        float x0, x1;
                        unpack2(x, x0, x1);
                        e0 = a4_ex2(x0);
                        e1 = a4_ex2(x1);
                    }
                    sr[c >> 1] = pack_bf16x2(e0, e1);
                }
                if (j > 0) {  // P buffer free: PV(j - 1) has read it (almost always true by now)
                    a4_wait(b_o_done, (j & 1) ^ 1);
                    tc_fence_after();
                }
                tmem_st16(tP, reinterpret_cast<uint32_t(&)[16]>(sr[0]));
                tmem_st_wait();
                tc_fence_before();
                a4_arrive(b_p_full);
            }
            // ---------------- epilogue: O / l -> out[b, t, h*d + :]; the two warps of a lane group take alternate 16-column chunks
            a4_wait(b_o_done, (nblk - 1) & 1);
            tc_fence_after();
            float l;
            {
                uint32_t r[16];
                tmem_ld16(tO + DN, r);
                tmem_ld_wait();
                l = __uint_as_float(r[0]);
            }
            const float inv_l = 1.0f / l;
            const int tq_row = q0 + t * 128 + row;
            if (hf == 0 && p.lse != nullptr && tq_row < p.tq) p.lse[(long long)bh * p.tq + tq_row] = m_used + log2f(l);
            const int b = bh / p.heads, h = bh - b * p.heads;
            __nv_bfloat16* orow = p.out + ((long long)b * p.tq + tq_row) * p.ld_out + h * p.head_dim;
#pragma unroll 1
            for (int c = hf * 16; c < DN; c += 32) {
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
    } else {
        if constexpr (Cfg::SOFTMAX_REGS == 96) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        else if constexpr (Cfg::SOFTMAX_REGS == 152) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ---------------- softmax / correction / epilogue of tile t: thread <-> query row
        const int t = warp >> 2;
        if (t < ntiles) {
            const int lg = warp & 3;
            const int row = lg * 32 + lane;
            const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
            const uint32_t tS = tmem_base + t * BKV + lane_addr;
            const uint32_t tP = tmem_base + P_BASE + t * P_STRIDE + lane_addr;
            const uint32_t tO = tmem_base + O_BASE + t * O_STRIDE + lane_addr;
            const uint32_t b_s_full = smem_u32(&s_full[t]), b_p_full = smem_u32(&p_full[t]);
            const uint32_t b_o_done = smem_u32(&o_done[t]), b_s_free = smem_u32(&s_free[t]);
            float m_used = -INFINITY;
            for (int j = 0; j < nblk; ++j) {
                uint32_t sr[BKV];
                a4_wait(b_s_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < BKV; c += 32) tmem_ld32(tS + c, reinterpret_cast<uint32_t(&)[32]>(sr[c]));
                tmem_ld_wait();
                if constexpr (DEC) {
                    tc_fence_before();
                    a4_arrive(b_s_free);  // the S columns may be overwritten by the next QK^T
                }
                const int valid = p.tk - j * BKV;  // keys of this block that exist (>= BKV except in the last block)
                if (valid < BKV) {
#pragma unroll
                    for (int i = 0; i < BKV; ++i)
                        if (i >= valid) sr[i] = 0xff800000u;  // -inf
                }
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains (latency)
#pragma unroll
                for (int i = 0; i < BKV; i += 8) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(sr[i + 2 * u]), __uint_as_float(sr[i + 2 * u + 1])));
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                if (j == 0) {
                    m_used = mx;
                } else {
                    const bool need = mx > m_used + 8.0f;
                    if (__any_sync(0xffffffffu, need)) {
                        // O and l are rescaled in place once PV(j - 1) and its row sums are complete (!DEC: S(j) came after
                        // them from the same issuing thread); PV(j) is not issued before this thread's p_full arrive below
                        if constexpr (DEC) {
                            a4_wait(b_o_done, (j & 1) ^ 1);
                            tc_fence_after();
                        }
                        const float m_new = fmaxf(m_used, mx);
                        const float alpha = a4_ex2(m_used - m_new);
                        m_used = m_new;
#pragma unroll 1
                        for (int c = 0; c < DN + 16; c += 16) {
                            uint32_t r[16];
                            tmem_ld16(tO + c, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            tmem_st16(tO + c, r);
                        }
                    }
                }
                // probabilities -> packed bf16 in registers (in place: sr[c/2] <- pack(p[c], p[c+1]))
                const uint64_t m2 = pack2(m_used, m_used);
#pragma unroll
                for (int c = 0; c < BKV; c += 2) {
                    const uint64_t x = sub2(pack2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), m2);
                    float e0, e1;
                    constexpr int kPolySlots[9] = {0x00, 0x08, 0x22, 0x2a, 0xaa, 0xab, 0xbb, 0xbf, 0xff};
                    if ((kPolySlots[POLY] >> ((c >> 1) & 7)) & 1) {
                        a4_ex2_poly2(x, e0, e1);
                    } else {
                        float x0, x1;
                        unpack2(x, x0, x1);
                        e0 = a4_ex2(x0);
                        e1 = a4_ex2(x1);
                    }
                    sr[c >> 1] = pack_bf16x2(e0, e1);
                }
                if constexpr (DEC) {
                    if (j > 0) {  // P buffer free: PV(j - 1) has read it (almost always true by now)
                        a4_wait(b_o_done, (j & 1) ^ 1);
                        tc_fence_after();
                    }
                }
                tmem_st32(tP, reinterpret_cast<uint32_t(&)[32]>(sr[0]));  // !DEC: over the first half of the S columns
                tmem_st_wait();
                tc_fence_before();
                a4_arrive(b_p_full);
            }
            // ---------------- epilogue: O / l -> out[b, t, h*d + :]
            a4_wait(b_o_done, DEC ? ((nblk - 1) & 1) : 0);
            tc_fence_after();
            float l;
            {
                uint32_t r[16];
                tmem_ld16(tO + DN, r);
                tmem_ld_wait();
                l = __uint_as_float(r[0]);
            }
            const float inv_l = 1.0f / l;
            const int tq_row = q0 + t * 128 + row;
            if (p.lse != nullptr && tq_row < p.tq) p.lse[(long long)bh * p.tq + tq_row] = m_used + log2f(l);
            const int b = bh / p.heads, h = bh - b * p.heads;
            __nv_bfloat16* orow = p.out + ((long long)b * p.tq + tq_row) * p.ld_out + h * p.head_dim;
#pragma unroll 1
            for (int c = 0; c < DN; c += 16) {
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

template <int NT, int BKV, int KVS, int POLY, int DK16, int DEC, int SPLIT = 0>
static int launch_attention4(const mobi_attn_args* a, AttnParams p, cudaStream_t stream) {
    const int d = a->head_dim;
    const long long BH = (long long)a->batch * a->heads;
    const long long smem = (long long)NT * p.nch * A4_CHUNK + 2ll * KVS * p.nch * BKV * 128 + A4_ONES_BYTES + 512 +
                           (SPLIT ? NT * 4 * 128 * 4 : 0) + 1024;
    const long long limit = 227 * 1024;
    MOBI_CHECK(smem <= limit, "mobi_attention: head_dim=%d needs %lld bytes of shared memory", d, smem);
    MOBI_CHECK(p.dk16 == DK16, "mobi_attention: head_dim=%d reached the kernel built for %d", d, DK16 * 16);
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
        MOBI_CUDA(cudaFuncSetAttribute(attention4_kernel<NT, BKV, KVS, POLY, DK16, DEC, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)limit));
        configured = true;
    }
    dim3 grid((a->tq + 128 * NT - 1) / (128 * NT), (unsigned)BH, 1);
    MOBI_CUDA(launch_pdl(attention4_kernel<NT, BKV, KVS, POLY, DK16, DEC, SPLIT>, grid, dim3(A4Cfg<NT, SPLIT>::THREADS), smem,
                         stream, tmQ, tmK, tmV, p));
    return 0;
}

// head_dim <= 48: four query tiles per CTA (O: 48 columns + 16 row-sum columns = the 64-column stride);
// head_dim <= 112: two tiles (O_STRIDE 128).  Everything else stays on attention3 (row sums in registers).
bool attention4_supports(int head_dim) { return head_dim <= 112; }

int attention4_dispatch(const mobi_attn_args* a, const AttnParams& p, cudaStream_t stream) {
    const int poly = (a->kernel >> 4) & 15;  // tuning hook: share of the exponentials on the FMA pipe (0 = default)
    const bool coupled = (a->kernel & 15) == 5;  // tuning hook / A-B arm: the first TMEM plan (NT = 4, P over S)
    switch (p.dk16) {
        case 1: return launch_attention4<3, 64, 8, 2, 1, 1>(a, p, stream);
        case 2: return launch_attention4<3, 64, 8, 2, 2, 1>(a, p, stream);
        case 3:  // head_dim 40: the level-0 attentions of the UNet
            if (coupled) return launch_attention4<4, 64, 6, 2, 3, 0>(a, p, stream);
            if ((a->kernel & 15) == 6) {  // tuning hook: split softmax (two warps per lane group)
                switch (poly) {
                    case 1: return launch_attention4<3, 64, 8, 0, 3, 1, 1>(a, p, stream);
                    case 2: return launch_attention4<3, 64, 8, 2, 3, 1, 1>(a, p, stream);
                    case 4: return launch_attention4<3, 64, 8, 4, 3, 1, 1>(a, p, stream);
                    default: return launch_attention4<3, 64, 8, 3, 3, 1, 1>(a, p, stream);
                }
            }
            switch (poly) {  // measured at the level-0 shape (32 rows): 1.41 / 1.39 / 1.25 / 1.22 / 1.27 ms
                case 1: return launch_attention4<3, 64, 8, 0, 3, 1>(a, p, stream);
                case 2: return launch_attention4<3, 64, 8, 1, 3, 1>(a, p, stream);
                case 3: return launch_attention4<3, 64, 8, 2, 3, 1>(a, p, stream);
                case 4: return launch_attention4<3, 64, 8, 4, 3, 1>(a, p, stream);
                default: return launch_attention4<3, 64, 8, 3, 3, 1>(a, p, stream);  // 3 of 8 pairs on the FMA pipe
            }
        case 4: return launch_attention4<2, 64, 3, 2, 4, 1>(a, p, stream);
        case 5:  // head_dim 80: level 1
            if (coupled) return launch_attention4<2, 64, 3, 2, 5, 0>(a, p, stream);
            if ((a->kernel & 15) == 6) return launch_attention4<2, 64, 3, 2, 5, 1, 1>(a, p, stream);
            return launch_attention4<2, 64, 3, 2, 5, 1>(a, p, stream);
        case 6: return launch_attention4<2, 64, 3, 2, 6, 1>(a, p, stream);
        default: return launch_attention4<2, 64, 3, 2, 7, 1>(a, p, stream);
    }
}

}  // namespace mobi
