// Kernels of the training step (LatentDiffusion.p_losses + autograd through the UNet, ldm/models/diffusion/ddpm.py:
// 1040-1058, 1177-1217): everything of the backward pass that is not a contraction.  The contractions (dgrad, wgrad,
// and the five products of the attention backward) run on the tcgen05 GEMM / implicit-conv kernels with transposed or
// flipped weight packs; this file holds the HBM-bound pieces around them:
//   transposes for the wgrad operands, LayerNorm / GroupNorm(+SiLU) / GEGLU backward, the softmax backward of the
//   attention (row statistics + dS, dS^T, P^T tiles), the 2-key context attention in q-space (forward + backward),
//   bias gradients, nearest-upsample / strided-conv adjoints, q_sample, the MSE loss gradient and AdamW.
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include <algorithm>
#include <cstdlib>
#include <cooperative_groups.h>

namespace mobi {

static inline unsigned bw_blocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

__device__ __forceinline__ float ld_any(const void* p, int is_f32, long long i) {
    return is_f32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// ------------------------------------------------------------------------------------------------
// Batched 2-D transpose: in [batch][rows, cols] (row stride ld_in) f32 or bf16 -> out bf16 [batch][cols, rows].
// 32x32 tiles through shared memory, both sides coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_kernel(const void* in, int in_f32, __nv_bfloat16* out, int rows, int cols, long long ld_in,
                 long long in_bs, long long ld_out, long long out_bs) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? ld_any(in, in_f32, b * in_bs + (long long)r * ld_in + c) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) out[b * out_bs + (long long)c * ld_out + r] = __float2bfloat16(tile[tx][i]);
    }
}

// bf16 -> bf16 fast path: 64 x 64 tiles moved as 32-bit (bf16x2) words on both sides.
__global__ void __launch_bounds__(256)
transpose_bf16x2_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int rows, int cols, long long ld_in,
                        long long in_bs, long long ld_out, long long out_bs) {
    __shared__ uint32_t t[64][33];  // [input row][input column pair]
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t* src = in + (b * in_bs) / 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + ty + 8 * i, c = c0 + 2 * tx;
        t[ty + 8 * i][tx] = (r < rows && c < cols) ? src[((long long)r * ld_in + c) >> 1] : 0u;
    }
    __syncthreads();
    uint32_t* dst = out + (b * out_bs) / 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int oc = ty + 8 * i;            // output row = input column c0 + oc
        const int r = r0 + 2 * tx;            // output column pair = input rows r, r + 1
        if (c0 + oc < cols && r < rows) {
            const uint32_t w0 = t[2 * tx][oc >> 1], w1 = t[2 * tx + 1][oc >> 1];
            const uint32_t v = (oc & 1) ? __byte_perm(w0, w1, 0x7632) : __byte_perm(w0, w1, 0x5410);
            dst[((long long)(c0 + oc) * ld_out + r) >> 1] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per token row (nn.LayerNorm, attention.py:213-223):
//   xh = (x - mean) * rstd ; g = dy * gamma ; dx = rstd * (g - mean(g) - xh * mean(g * xh))
//   dgamma += sum_rows dy * xh ; dbeta += sum_rows dy      (only for the trainable adapter norms)
// Rows are addressed through the same segment gather as the forward (camera / lidar rows of the interleaved batch).
// ------------------------------------------------------------------------------------------------
constexpr int LNB_MAXV = 40;  // C <= 1280

__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const void* __restrict__ dy, int dy_f32,
              float* dx, float* dgamma, float* dbeta, long long rows, int C, long long seg, long long seg_stride,
              long long seg_offset, float eps, int accumulate) {
    extern __shared__ float lnb_smem[];  // [2][C] when dgamma
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    if (dgamma) {
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) lnb_smem[c] = 0.f;
        __syncthreads();
    }
    const float inv_c = 1.0f / C;
    for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
        const long long src = seg > 0 ? (row / seg) * seg_stride + seg_offset + row % seg : row;
        const float* xr = x + src * C;
        float xv[LNB_MAXV], gv[LNB_MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int c = lane + 32 * i;
            xv[i] = c < C ? xr[c] : 0.f;
            s += xv[i];
        }
        const float mean = warp_sum(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int c = lane + 32 * i;
            const float d = c < C ? xv[i] - mean : 0.f;
            xv[i] = d;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int c = lane + 32 * i;
            float g = 0.f;
            if (c < C) {
                const float dyv = ld_any(dy, dy_f32, row * C + c);
                xv[i] *= rstd;  // xhat
                g = dyv * (gamma ? gamma[c] : 1.0f);
                if (dgamma) {
                    atomicAdd(&lnb_smem[c], dyv * xv[i]);
                    atomicAdd(&lnb_smem[C + c], dyv);
                }
            }
            gv[i] = g;
            sg += g;
            sgx += g * xv[i];
        }
        const float mg = warp_sum(sg) * inv_c, mgx = warp_sum(sgx) * inv_c;
        float* dr = dx + src * C;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                const float v = rstd * (gv[i] - mg - xv[i] * mgx);
                dr[c] = accumulate ? dr[c] + v : v;
            }
        }
    }
    if (dgamma) {
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            atomicAdd(dgamma + c, lnb_smem[c]);
            atomicAdd(dbeta + c, lnb_smem[C + c]);
        }
    }
}

// Width- and dtype-specialised LayerNorm backward: a lane owns channel PAIRS (c, c + 1) = 2 * lane + 64 * j, so x / dx move
// as 8-byte and bf16 dy as 4-byte words (256 / 128 contiguous bytes per warp load); NP = pairs per lane (5 / 10 / 20 for
// C <= 320 / 640 / 1280); the dgamma / dbeta partial sums of all rows a warp visits stay in registers (NP <= 5: C <= 320, where most token rows live) and
// reach shared memory once per warp; the dx read of the accumulate path is issued together with the x / dy loads.
template <int NP, bool DY_F32>
__global__ void __launch_bounds__(256)
ln_bwd2_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const void* __restrict__ dy,
               float* dx, float* dgamma, float* dbeta, long long rows, int C, long long seg, long long seg_stride,
               long long seg_offset, float eps, int accumulate) {
    extern __shared__ float lnb_smem[];  // [2][C] when dgamma
    constexpr bool REG_ACC = NP <= 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    if (dgamma) {
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) lnb_smem[c] = 0.f;
        __syncthreads();
    }
    float2 ag[REG_ACC ? NP : 1], ab[REG_ACC ? NP : 1];
#pragma unroll
    for (int i = 0; i < (REG_ACC ? NP : 1); ++i) ag[i] = ab[i] = make_float2(0.f, 0.f);
    const float inv_c = 1.0f / C;
    for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
        const long long src = seg > 0 ? (row / seg) * seg_stride + seg_offset + row % seg : row;
        const float2* xr = reinterpret_cast<const float2*>(x + src * C);
        float2* dr = reinterpret_cast<float2*>(dx + src * C);
        float2 xv[NP], gv[NP], acc[NP];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int pidx = lane + 32 * j;  // pair index: channels 2 * pidx, 2 * pidx + 1
            const bool ok = 2 * pidx < C;
            xv[j] = ok ? xr[pidx] : make_float2(0.f, 0.f);
            if (DY_F32) {
                gv[j] = ok ? reinterpret_cast<const float2*>(reinterpret_cast<const float*>(dy) + row * C)[pidx] : make_float2(0.f, 0.f);
            } else {
                gv[j] = ok ? __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(
                                 reinterpret_cast<const __nv_bfloat16*>(dy) + row * C)[pidx])
                           : make_float2(0.f, 0.f);
            }
            acc[j] = (accumulate && ok) ? dr[pidx] : make_float2(0.f, 0.f);
            s += xv[j].x + xv[j].y;
        }
        const float mean = warp_sum(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const bool ok = 2 * (lane + 32 * j) < C;
            xv[j].x = ok ? xv[j].x - mean : 0.f;
            xv[j].y = ok ? xv[j].y - mean : 0.f;
            q += xv[j].x * xv[j].x + xv[j].y * xv[j].y;
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int pidx = lane + 32 * j;
            const bool ok = 2 * pidx < C;
            const float2 dyv = gv[j];
            xv[j].x *= rstd;  // xhat (0 beyond C)
            xv[j].y *= rstd;
            if (dgamma) {
                if (REG_ACC) {
                    ag[j].x = fmaf(dyv.x, xv[j].x, ag[j].x);
                    ag[j].y = fmaf(dyv.y, xv[j].y, ag[j].y);
                    ab[j].x += dyv.x;
                    ab[j].y += dyv.y;
                } else if (ok) {
                    atomicAdd(&lnb_smem[2 * pidx], dyv.x * xv[j].x);
                    atomicAdd(&lnb_smem[2 * pidx + 1], dyv.y * xv[j].y);
                    atomicAdd(&lnb_smem[C + 2 * pidx], dyv.x);
                    atomicAdd(&lnb_smem[C + 2 * pidx + 1], dyv.y);
                }
            }
            const float2 gm = (gamma && ok) ? reinterpret_cast<const float2*>(gamma)[pidx] : make_float2(1.f, 1.f);
            gv[j].x = dyv.x * gm.x;
            gv[j].y = dyv.y * gm.y;
            sg += gv[j].x + gv[j].y;
            sgx += gv[j].x * xv[j].x + gv[j].y * xv[j].y;
        }
        const float mg = warp_sum(sg) * inv_c, mgx = warp_sum(sgx) * inv_c;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int pidx = lane + 32 * j;
            if (2 * pidx < C)
                dr[pidx] = make_float2(acc[j].x + rstd * (gv[j].x - mg - xv[j].x * mgx),
                                       acc[j].y + rstd * (gv[j].y - mg - xv[j].y * mgx));
        }
    }
    if (dgamma) {
        if (REG_ACC) {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int pidx = lane + 32 * j;
                if (2 * pidx < C) {
                    atomicAdd(&lnb_smem[2 * pidx], ag[j].x);
                    atomicAdd(&lnb_smem[2 * pidx + 1], ag[j].y);
                    atomicAdd(&lnb_smem[C + 2 * pidx], ab[j].x);
                    atomicAdd(&lnb_smem[C + 2 * pidx + 1], ab[j].y);
                }
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            atomicAdd(dgamma + c, lnb_smem[c]);
            atomicAdd(dbeta + c, lnb_smem[C + c]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) backward over NHWC, one 8-CTA thread-block cluster per (image, group) (each CTA owns a pixel slab,
// the two reductions go through distributed shared memory), three passes over the slab (which stays in L1/L2 between
// passes): statistics; S1 = sum dz, S2 = sum dz * xh with dz = dy * silu'(y) * gamma; then
//   dx = rstd * (dz - (S1 + xh * S2) / m) + dres
// The input may be the channel concatenation of two tensors (skip connections, openaimodel.py:892): dx is split.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

constexpr int GNB_CLUSTER = 8;  // CTAs per (image, group): a thread-block cluster that reduces through DSMEM

// Sum of (a, b) over all threads of all CTAs of the cluster; every thread gets the totals.
__device__ __forceinline__ void cluster_sum2(float& a, float& b, float* red, float* slot) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const float ta = block_sum(a, red);
    const float tb = block_sum(b, red);
    if (threadIdx.x == 0) {
        slot[0] = ta;
        slot[1] = tb;
    }
    cluster.sync();
    float sa = 0.f, sb = 0.f;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
        const float* remote = cluster.map_shared_rank(slot, r);
        sa += remote[0];
        sb += remote[1];
    }
    cluster.sync();  // nobody overwrites its slot before every CTA has read it
    a = sa;
    b = sb;
}

__global__ void __cluster_dims__(GNB_CLUSTER, 1, 1) __launch_bounds__(256)
gn_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ gamma,
              const float* __restrict__ beta, const void* __restrict__ dy, int dy_f32, const float* __restrict__ dres,
              float* dx1, float* dx2, int hw, int c1, int c2, int groups, int silu, float eps) {
    __shared__ float red[8];
    __shared__ float slot[2];
    const int C = c1 + c2;
    const int cpg = C / groups;
    const int n = blockIdx.y, g = blockIdx.x / GNB_CLUSTER, part = blockIdx.x % GNB_CLUSTER;
    const int cbase = g * cpg;
    const long long img = (long long)n * hw;
    const int pix_per = (hw + GNB_CLUSTER - 1) / GNB_CLUSTER;
    const int pbeg = min(hw, part * pix_per), pend = min(hw, pbeg + pix_per);
    const int total = (pend - pbeg) * cpg;
    auto load_x = [&](int p, int c) -> float {
        return c < c1 ? x1[(img + p) * c1 + c] : x2[(img + p) * c2 + (c - c1)];
    };
    float s = 0.f, q = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int pl = i / cpg, c = cbase + i - pl * cpg;
        const float v = load_x(pbeg + pl, c);
        s += v;
        q += v * v;
    }
    cluster_sum2(s, q, red, slot);
    const float m = (float)hw * cpg;
    const float mean = s / m;
    const float var = fmaxf(q / m - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int pl = i / cpg, c = cbase + i - pl * cpg;
        const int p = pbeg + pl;
        const float xh = (load_x(p, c) - mean) * rstd;
        float dz = ld_any(dy, dy_f32, (img + p) * C + c);
        if (silu) {
            const float y = xh * gamma[c] + beta[c];
            const float sg = 1.0f / (1.0f + __expf(-y));
            dz *= sg * (1.0f + y * (1.0f - sg));
        }
        dz *= gamma[c];
        s1 += dz;
        s2 += dz * xh;
    }
    cluster_sum2(s1, s2, red, slot);
    const float S1 = s1 / m, S2 = s2 / m;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int pl = i / cpg, c = cbase + i - pl * cpg;
        const int p = pbeg + pl;
        const float xh = (load_x(p, c) - mean) * rstd;
        float dz = ld_any(dy, dy_f32, (img + p) * C + c);
        if (silu) {
            const float y = xh * gamma[c] + beta[c];
            const float sg = 1.0f / (1.0f + __expf(-y));
            dz *= sg * (1.0f + y * (1.0f - sg));
        }
        dz *= gamma[c];
        float v = rstd * (dz - S1 - xh * S2);
        if (dres) v += dres[(img + p) * C + c];
        if (c < c1) dx1[(img + p) * c1 + c] = v;
        else dx2[(img + p) * c2 + (c - c1)] = v;
    }
}

// Channel-pair version (every UNet shape: channels per group and c1 are even): a thread owns ONE channel pair of the group
// for the whole kernel (gamma / beta / source tensor fixed in registers, no per-element division or dtype branch) and
// walks pixels, four in flight, with 8-byte accesses.  Same three passes, same cluster reductions.
template <bool DY_F32>
__global__ void __cluster_dims__(GNB_CLUSTER, 1, 1) __launch_bounds__(256)
gn_bwd2_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ gamma,
               const float* __restrict__ beta, const void* __restrict__ dy, const float* __restrict__ dres,
               float* dx1, float* dx2, int hw, int c1, int c2, int groups, int silu, float eps) {
    __shared__ float red[8];
    __shared__ float slot[2];
    const int C = c1 + c2;
    const int cpg = C / groups;
    const int npairs = cpg >> 1;
    const int n = blockIdx.y, g = blockIdx.x / GNB_CLUSTER, part = blockIdx.x % GNB_CLUSTER;
    const int pv = threadIdx.x % npairs, lane_p = threadIdx.x / npairs;
    const int lanes = blockDim.x / npairs;
    const bool active = lane_p < lanes;
    const int c = g * cpg + 2 * pv;  // first channel of this thread's pair
    const long long img = (long long)n * hw;
    const int pix_per = (hw + GNB_CLUSTER - 1) / GNB_CLUSTER;
    const int pbeg = min(hw, part * pix_per), pend = min(hw, pbeg + pix_per);
    // source / destination of this pair: x1 | x2 are the two halves of a channel concatenation
    const bool first = c < c1;
    const float* xs = first ? x1 + c : x2 + (c - c1);
    float* dxs = first ? dx1 + c : dx2 + (c - c1);
    const int cs = first ? c1 : c2;
    const float2 gm = *reinterpret_cast<const float2*>(gamma + c);
    const float2 bt = *reinterpret_cast<const float2*>(beta + c);
    auto ld_x = [&](int p) { return *reinterpret_cast<const float2*>(xs + (img + p) * cs); };
    auto ld_dy = [&](int p) {
        if (DY_F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(dy) + (img + p) * C + c);
        return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (img + p) * C + c));
    };
    constexpr int U = 4;
    float s = 0.f, q = 0.f;
    if (active) {
        for (int p = pbeg + lane_p; p < pend; p += U * lanes) {
            float2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = (p + u * lanes < pend) ? ld_x(p + u * lanes) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                s += v[u].x + v[u].y;
                q += v[u].x * v[u].x + v[u].y * v[u].y;
            }
        }
    }
    cluster_sum2(s, q, red, slot);
    const float m = (float)hw * cpg;
    const float mean = s / m;
    const float var = fmaxf(q / m - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    // dz = dy * silu'(y) * gamma for one element
    auto dz_of = [&](float xh, float dyv, float gmc, float btc) {
        float dz = dyv;
        if (silu) {
            const float y = xh * gmc + btc;
            const float sg = 1.0f / (1.0f + __expf(-y));
            dz *= sg * (1.0f + y * (1.0f - sg));
        }
        return dz * gmc;
    };
    float s1 = 0.f, s2 = 0.f;
    if (active) {
        for (int p = pbeg + lane_p; p < pend; p += U * lanes) {
            float2 v[U], d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = p + u * lanes < pend;
                v[u] = ok ? ld_x(p + u * lanes) : make_float2(mean, mean);
                d[u] = ok ? ld_dy(p + u * lanes) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float xh0 = (v[u].x - mean) * rstd, xh1 = (v[u].y - mean) * rstd;
                const float dz0 = dz_of(xh0, d[u].x, gm.x, bt.x), dz1 = dz_of(xh1, d[u].y, gm.y, bt.y);
                s1 += dz0 + dz1;
                s2 += dz0 * xh0 + dz1 * xh1;
            }
        }
    }
    cluster_sum2(s1, s2, red, slot);
    const float S1 = s1 / m, S2 = s2 / m;
    if (!active) return;
    for (int p = pbeg + lane_p; p < pend; p += U * lanes) {
        float2 v[U], d[U], r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = p + u * lanes < pend;
            v[u] = ok ? ld_x(p + u * lanes) : make_float2(mean, mean);
            d[u] = ok ? ld_dy(p + u * lanes) : make_float2(0.f, 0.f);
            r[u] = (ok && dres) ? *reinterpret_cast<const float2*>(dres + (img + p + u * lanes) * C + c) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * lanes >= pend) break;
            const float xh0 = (v[u].x - mean) * rstd, xh1 = (v[u].y - mean) * rstd;
            const float dz0 = dz_of(xh0, d[u].x, gm.x, bt.x), dz1 = dz_of(xh1, d[u].y, gm.y, bt.y);
            *reinterpret_cast<float2*>(dxs + (img + p + u * lanes) * cs) =
                make_float2(rstd * (dz0 - S1 - xh0 * S2) + r[u].x, rstd * (dz1 - S1 - xh1 * S2) + r[u].y);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GEGLU (attention.py:38-45) on (value, gate) column pairs, exact erf GELU (F.gelu default).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__global__ void geglu_fwd_kernel(const __nv_bfloat162* __restrict__ g, __nv_bfloat16* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 vg = __bfloat1622float2(g[i]);
    out[i] = __float2bfloat16(vg.x * gelu_erf(vg.y));
}

__global__ void geglu_bwd_kernel(const __nv_bfloat162* __restrict__ g, const __nv_bfloat16* __restrict__ dh,
                                 __nv_bfloat162* __restrict__ dg, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 vg = __bfloat1622float2(g[i]);
    const float d = __bfloat162float(dh[i]);
    dg[i] = __floats2bfloat162_rn(d * gelu_erf(vg.y), d * vg.x * gelu_erf_grad(vg.y));
}

// ------------------------------------------------------------------------------------------------
// Softmax backward of one attention (CrossAttention.forward, attention.py:181-190) from materialised tiles:
//   S  = q' k^T  (log2 domain: the scale * log2(e) is folded into q'),  dP = dO V^T     (both f32 [Tq, Tk])
//   P = 2^(S - rowmax) / rowsum ;  Delta = rowsum(P * dP) ;  dS = dscale * P * (dP - Delta)
// Pass 1 (one warp per row, online): rowmax, rowsum, Delta.  Pass 2 (32x32 tiles): dS row-major and, through a
// shared-memory transpose, dS^T and P^T (the K-major operands of dK = dS^T q' and dV = P^T dO).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_bwd_stats_kernel(const float* __restrict__ S, const float* __restrict__ dP, float* __restrict__ stats, int tq,
                      int tk, long long batch_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= tq) return;
    const long long b = blockIdx.y;
    const float* s = S + b * batch_stride + (long long)row * tk;
    const float* d = dP + b * batch_stride + (long long)row * tk;
    float m = -INFINITY, l = 0.f, acc = 0.f;
    for (int j = lane; j < tk; j += 32) {
        const float sv = s[j];
        if (sv > m) {
            const float r = exp2f(m - sv);
            l *= r;
            acc *= r;
            m = sv;
        }
        const float e = exp2f(sv - m);
        l += e;
        acc += e * d[j];
    }
    const float M = warp_max(m);
    const float r = (m == -INFINITY) ? 0.f : exp2f(m - M);
    l = warp_sum(l * r);
    acc = warp_sum(acc * r);
    if (lane == 0) {
        float* o = stats + (b * tq + row) * 3;
        o[0] = M;
        o[1] = 1.0f / l;
        o[2] = acc / l;
    }
}

// stats[b, row] = (lse, 1, Delta) with Delta = <dO_row, O_row> over the head's d columns: one warp per (head, row)
__global__ void __launch_bounds__(256)
attn_bwd_stats_from_lse_kernel(const float* __restrict__ lse, const __nv_bfloat16* __restrict__ o,
                               const __nv_bfloat16* __restrict__ d_o, float* __restrict__ stats, int tq, long long ld,
                               int head_dim) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= tq) return;
    const long long b = blockIdx.y;
    const long long base = (long long)row * ld + b * head_dim;
    float acc = 0.f;
    for (int j = lane; j < head_dim; j += 32) acc += __bfloat162float(o[base + j]) * __bfloat162float(d_o[base + j]);
    acc = warp_sum(acc);
    if (lane == 0) {
        float* st = stats + (b * tq + row) * 3;
        st[0] = lse[b * tq + row];
        st[1] = 1.0f;
        st[2] = acc;
    }
}

__global__ void __launch_bounds__(256)
attn_bwd_apply_kernel(const float* __restrict__ S, const float* __restrict__ dP, const float* __restrict__ stats,
                      __nv_bfloat16* __restrict__ dS, __nv_bfloat16* __restrict__ dSt, __nv_bfloat16* __restrict__ Pt,
                      int tq, int tk, float dscale, long long batch_stride) {
    __shared__ float tp[32][33];
    __shared__ float td[32][33];
    const long long b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long base = b * batch_stride;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        float p = 0.f, ds = 0.f;
        if (r < tq && c < tk) {
            const float* st = stats + (b * tq + r) * 3;
            const long long idx = base + (long long)r * tk + c;
            p = exp2f(S[idx] - st[0]) * st[1];
            ds = dscale * p * (dP[idx] - st[2]);
            dS[idx] = __float2bfloat16(ds);
        }
        tp[i][tx] = p;
        td[i][tx] = ds;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < tk && r < tq) {
            const long long idx = base + (long long)c * tq + r;
            dSt[idx] = __float2bfloat16(td[tx][i]);
            Pt[idx] = __float2bfloat16(tp[tx][i]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cond_adapter_attn (attention.py:237-243) in q-space for the training step: `keys` (1..4) context tokens per batch
// row, `heads` heads of D = C / heads.  One CTA per (token chunk of 128, head, batch row).
//   forward : s_j = <q_t, k_j> (scale folded into q), p = softmax_j, o_t = sum_j p_j v_j
//   backward: dp_j = <do_t, v_j>, ds_j = p_j (dp_j - sum p dp), dq_t = sum_j ds_j k_j,
//             dk_j += sum_t ds_j q_t, dv_j += sum_t p_j do_t   (block reduction, then one atomicAdd per element)
// ------------------------------------------------------------------------------------------------
constexpr int CA_MAXK = 4;
constexpr int CA_TOK = 128;

template <bool BWD>
__global__ void __launch_bounds__(CA_TOK)
ctx_attn_qspace_kernel(const __nv_bfloat16* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                       __nv_bfloat16* __restrict__ dq, float* dk, float* dv, int tokens, int C, int heads, int keys) {
    extern __shared__ float ca_smem[];  // k[keys][D], v[keys][D], p[keys][CA_TOK], ds[keys][CA_TOK]
    const int D = C / heads;
    const int b = blockIdx.z, h = blockIdx.y;
    const int t = blockIdx.x * CA_TOK + threadIdx.x;
    float* sk = ca_smem;
    float* sv = sk + keys * D;
    float* sp = sv + keys * D;
    float* sds = sp + keys * CA_TOK;
    for (int i = threadIdx.x; i < keys * D; i += blockDim.x) {
        const int j = i / D, dd = i - j * D;
        sk[i] = k[((long long)b * keys + j) * C + h * D + dd];
        sv[i] = v[((long long)b * keys + j) * C + h * D + dd];
    }
    __syncthreads();
    const bool ok = t < tokens;
    const long long row = ((long long)b * tokens + t) * C + h * D;
    float p[CA_MAXK], ds[CA_MAXK];
#pragma unroll
    for (int j = 0; j < CA_MAXK; ++j) p[j] = ds[j] = 0.f;
    if (ok) {
        float s[CA_MAXK];
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j) s[j] = 0.f;
        for (int dd = 0; dd < D; ++dd) {
            const float qv = __bfloat162float(q[row + dd]);
#pragma unroll
            for (int j = 0; j < CA_MAXK; ++j)
                if (j < keys) s[j] += qv * sk[j * D + dd];
        }
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j)
            if (j < keys) m = fmaxf(m, s[j]);
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j)
            if (j < keys) {
                p[j] = __expf(s[j] - m);
                l += p[j];
            }
        const float inv = 1.0f / l;
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j) p[j] *= inv;
        if (!BWD) {
            for (int dd = 0; dd < D; ++dd) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < CA_MAXK; ++j)
                    if (j < keys) acc += p[j] * sv[j * D + dd];
                o[row + dd] = __float2bfloat16(acc);
            }
        } else {
            float dp[CA_MAXK];
#pragma unroll
            for (int j = 0; j < CA_MAXK; ++j) dp[j] = 0.f;
            for (int dd = 0; dd < D; ++dd) {
                const float dov = __bfloat162float(d_o[row + dd]);
#pragma unroll
                for (int j = 0; j < CA_MAXK; ++j)
                    if (j < keys) dp[j] += dov * sv[j * D + dd];
            }
            float delta = 0.f;
#pragma unroll
            for (int j = 0; j < CA_MAXK; ++j) delta += p[j] * dp[j];
#pragma unroll
            for (int j = 0; j < CA_MAXK; ++j) ds[j] = p[j] * (dp[j] - delta);
            for (int dd = 0; dd < D; ++dd) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < CA_MAXK; ++j)
                    if (j < keys) acc += ds[j] * sk[j * D + dd];
                dq[row + dd] = __float2bfloat16(acc);
            }
        }
    }
    if (BWD) {
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j)
            if (j < keys) {
                sp[j * CA_TOK + threadIdx.x] = p[j];
                sds[j * CA_TOK + threadIdx.x] = ds[j];
            }
        __syncthreads();
        const int t0 = blockIdx.x * CA_TOK;
        const int nt = min(CA_TOK, tokens - t0);
        for (int i = threadIdx.x; i < keys * D; i += blockDim.x) {
            const int j = i / D, dd = i - j * D;
            float ak = 0.f, av = 0.f;
            const long long col = ((long long)b * tokens + t0) * C + h * D + dd;
            for (int tt = 0; tt < nt; ++tt) {
                ak += sds[j * CA_TOK + tt] * __bfloat162float(q[col + (long long)tt * C]);
                av += sp[j * CA_TOK + tt] * __bfloat162float(d_o[col + (long long)tt * C]);
            }
            atomicAdd(dk + ((long long)b * keys + j) * C + h * D + dd, ak);
            atomicAdd(dv + ((long long)b * keys + j) * C + h * D + dd, av);
        }
    }
}

// Vectorised version for D % 8 == 0 (every level of the UNet: D = 40 / 80 / 160): a thread owns one token of one head and
// moves its q / dO / o / dq segments as 16-byte words (D / 8 of them) held in registers; the backward also parks q and dO
// of the 128 tokens in shared memory so that the dk / dv block sums read them there instead of walking global columns.
template <bool BWD, int DV>  // DV = D / 8
__global__ void __launch_bounds__(CA_TOK)
ctx_attn_qspace_vec_kernel(const __nv_bfloat16* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                           __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                           __nv_bfloat16* __restrict__ dq, float* dk, float* dv, int tokens, int C, int heads, int keys) {
    extern __shared__ __align__(16) float ca_smem[];  // k[keys][D], v[keys][D], p[keys][128], ds[keys][128], q/dO tiles (bf16)
    constexpr int D = DV * 8;
    const int b = blockIdx.z, h = blockIdx.y;
    const int t = blockIdx.x * CA_TOK + threadIdx.x;
    float* sk = ca_smem;
    float* sv = sk + keys * D;
    float* sp = sv + keys * D;
    float* sds = sp + keys * CA_TOK;
    // bf16 tiles [128][D + 8] (row pitch padded by 16 bytes: conflict-free 16-byte row writes and column reads)
    constexpr int PITCH = D + 8;
    __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(sds + keys * CA_TOK);
    __nv_bfloat16* sdo = sq + CA_TOK * PITCH;
    for (int i = threadIdx.x; i < keys * D; i += blockDim.x) {
        const int j = i / D, dd = i - j * D;
        sk[i] = k[((long long)b * keys + j) * C + h * D + dd];
        sv[i] = v[((long long)b * keys + j) * C + h * D + dd];
    }
    __syncthreads();
    const bool ok = t < tokens;
    const long long row = ((long long)b * tokens + t) * C + h * D;
    float p[CA_MAXK], ds[CA_MAXK];
#pragma unroll
    for (int j = 0; j < CA_MAXK; ++j) p[j] = ds[j] = 0.f;
    uint4 qv[DV], dov[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) qv[i] = dov[i] = make_uint4(0u, 0u, 0u, 0u);
    if (ok) {
#pragma unroll
        for (int i = 0; i < DV; ++i) {
            qv[i] = __ldg(reinterpret_cast<const uint4*>(q + row) + i);
            if (BWD) dov[i] = __ldg(reinterpret_cast<const uint4*>(d_o + row) + i);
        }
        float s[CA_MAXK], dp[CA_MAXK];
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j) s[j] = dp[j] = 0.f;
#pragma unroll
        for (int i = 0; i < DV; ++i) {
            const uint32_t qw[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
            const uint32_t dw[4] = {dov[i].x, dov[i].y, dov[i].z, dov[i].w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float q0 = __uint_as_float(qw[w] << 16), q1 = __uint_as_float(qw[w] & 0xffff0000u);
                const float d0 = __uint_as_float(dw[w] << 16), d1 = __uint_as_float(dw[w] & 0xffff0000u);
                const int dd = i * 8 + w * 2;
#pragma unroll
                for (int j = 0; j < CA_MAXK; ++j)
                    if (j < keys) {
                        s[j] = fmaf(q0, sk[j * D + dd], fmaf(q1, sk[j * D + dd + 1], s[j]));
                        if (BWD) dp[j] = fmaf(d0, sv[j * D + dd], fmaf(d1, sv[j * D + dd + 1], dp[j]));
                    }
            }
        }
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j)
            if (j < keys) m = fmaxf(m, s[j]);
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j)
            if (j < keys) {
                p[j] = __expf(s[j] - m);
                l += p[j];
            }
        const float inv = 1.0f / l;
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j) p[j] *= inv;
        if (BWD) {
            float delta = 0.f;
#pragma unroll
            for (int j = 0; j < CA_MAXK; ++j) delta += p[j] * dp[j];
#pragma unroll
            for (int j = 0; j < CA_MAXK; ++j) ds[j] = p[j] * (dp[j] - delta);
        }
        // o = sum_j p_j v_j (forward) / dq = sum_j ds_j k_j (backward), 8 channels per 16-byte store
        const float* coef = BWD ? ds : p;
        const float* tab = BWD ? sk : sv;
        __nv_bfloat16* dst = (BWD ? dq : o) + row;
#pragma unroll
        for (int i = 0; i < DV; ++i) {
            uint32_t wds[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int dd = i * 8 + w * 2;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int j = 0; j < CA_MAXK; ++j)
                    if (j < keys) {
                        a0 = fmaf(coef[j], tab[j * D + dd], a0);
                        a1 = fmaf(coef[j], tab[j * D + dd + 1], a1);
                    }
                wds[w] = pack_bf16x2(a0, a1);
            }
            reinterpret_cast<uint4*>(dst)[i] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        }
    }
    if (BWD) {
#pragma unroll
        for (int j = 0; j < CA_MAXK; ++j)
            if (j < keys) {
                sp[j * CA_TOK + threadIdx.x] = p[j];    // zero for tokens beyond the end
                sds[j * CA_TOK + threadIdx.x] = ds[j];
            }
#pragma unroll
        for (int i = 0; i < DV; ++i) {
            *reinterpret_cast<uint4*>(sq + threadIdx.x * PITCH + i * 8) = qv[i];
            *reinterpret_cast<uint4*>(sdo + threadIdx.x * PITCH + i * 8) = dov[i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < keys * D; i += blockDim.x) {
            const int j = i / D, dd = i - j * D;
            float ak = 0.f, av = 0.f;
#pragma unroll 8
            for (int tt = 0; tt < CA_TOK; ++tt) {
                ak = fmaf(sds[j * CA_TOK + tt], __bfloat162float(sq[tt * PITCH + dd]), ak);
                av = fmaf(sp[j * CA_TOK + tt], __bfloat162float(sdo[tt * PITCH + dd]), av);
            }
            atomicAdd(dk + ((long long)b * keys + j) * C + h * D + dd, ak);
            atomicAdd(dv + ((long long)b * keys + j) * C + h * D + dd, av);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// out[g, c] += sum over the rows of group g of x[r, c]   (bias gradients; per-batch-row sums)
// grid (ceil(cols / 32), row chunks); block 32 x 8.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, int is_f32, long long rows, int cols, long long ld, long long rows_per_group,
              int chunk, float* out) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const long long r0 = (long long)blockIdx.y * chunk;
    const long long r1 = min(rows, r0 + chunk);
    float acc = 0.f;
    if (c < cols)
        for (long long r = r0 + ty; r < r1; r += 8) acc += ld_any(x, is_f32, r * ld + c);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        atomicAdd(out + (r0 / rows_per_group) * cols + c, t);
    }
}

// Column-pair version: a warp reads 64 consecutive column pairs of a row (128 / 256 contiguous bytes for bf16 / f32), a
// block covers 128 columns x `chunk` rows with 4 row lanes and eight rows in flight per thread.
template <bool F32>
__global__ void __launch_bounds__(256)
colsum2_kernel(const void* __restrict__ x, long long rows, int cols, long long ld, long long rows_per_group, int chunk,
               float* out) {
    __shared__ float2 red[4][64];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int c = blockIdx.x * 128 + 2 * tx;
    const long long r0 = (long long)blockIdx.y * chunk;
    const long long r1 = min(rows, r0 + chunk);
    float2 acc = make_float2(0.f, 0.f);
    if (c < cols) {
        auto ld2 = [&](long long r) {
            if (F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(x) + r * ld + c);
            return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(x) + r * ld + c));
        };
        for (long long r = r0 + ty; r < r1; r += 32) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (r + 4 * u < r1) ? ld2(r + 4 * u) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc.x += v[u].x;
                acc.y += v[u].y;
            }
        }
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
        const float2 a = red[0][tx], b = red[1][tx], d = red[2][tx], e = red[3][tx];
        float* dst = out + (r0 / rows_per_group) * cols + c;
        atomicAdd(dst, (a.x + b.x) + (d.x + e.x));
        atomicAdd(dst + 1, (a.y + b.y) + (d.y + e.y));
    }
}

// out[n, k] += sum_m A[m, n] * B[m, k] for tiny m (the context-token side of the adapters: m = batch * keys)
__global__ void wgrad_small_kernel(const float* __restrict__ A, const float* __restrict__ B, float* out, int m, int n,
                                   int k, long long lda, long long ldb, long long ldo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * k) return;
    const int nn = (int)(i / k), kk = (int)(i - (long long)nn * k);
    float acc = 0.f;
    for (int mm = 0; mm < m; ++mm) acc += A[mm * lda + nn] * B[mm * ldb + kk];
    out[nn * ldo + kk] += acc;
}

// ------------------------------------------------------------------------------------------------
// Adjoints of the resampling layers.  zero_insert2x: dy [N,h,w,C] -> z [N,2h,2w,C] bf16 with dy at the even positions
// (the strided conv3x3 of Downsample, openaimodel.py:151-153, becomes a stride-1 conv of z with the flipped filter);
// sum2x2: adjoint of F.interpolate(scale_factor=2, mode="nearest") (openaimodel.py:116).
// ------------------------------------------------------------------------------------------------
__global__ void zero_insert2x_kernel(const void* __restrict__ dy, int is_f32, __nv_bfloat16* __restrict__ z, int n, int h,
                                     int w, int c) {
    const long long total = (long long)n * 4 * h * w * c;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cc = (int)(i % c);
    long long p = i / c;
    const int x = (int)(p % (2 * w));
    p /= 2 * w;
    const int y = (int)(p % (2 * h));
    const int img = (int)(p / (2 * h));
    float v = 0.f;
    if (((x | y) & 1) == 0) v = ld_any(dy, is_f32, (((long long)img * h + (y >> 1)) * w + (x >> 1)) * c + cc);
    z[i] = __float2bfloat16(v);
}

__global__ void sum2x2_kernel(const float* __restrict__ d, float* __restrict__ out, int n, int h, int w, int c) {
    const long long total = (long long)n * h * w * c;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cc = (int)(i % c);
    long long p = i / c;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const int img = (int)(p / h);
    const long long row = (long long)2 * w * c;
    const float* s = d + (((long long)img * 2 * h + 2 * y) * 2 * w + 2 * x) * c + cc;
    out[i] = s[0] + s[c] + s[row] + s[row + c];
}

// ------------------------------------------------------------------------------------------------
// q_sample on the first `c_noised` channels (ddpm.py:284-287, 1178-1182): NCHW f32.
// ------------------------------------------------------------------------------------------------
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac,
                                const long long* __restrict__ t, float* __restrict__ out, int B, int c_total,
                                int c_noised, int hw) {
    const long long total = (long long)B * c_total * hw;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / ((long long)c_total * hw));
    const int rem = (int)(i - (long long)b * c_total * hw);
    const int c = rem / hw;
    float v = x0[i];
    if (c < c_noised) {
        const long long tt = t[b];
        v = sqrt_ac[tt] * v + sqrt_1mac[tt] * noise[((long long)b * c_noised + c) * hw + (rem - c * hw)];
    }
    out[i] = v;
}

// loss_sum += sum (pred - target)^2 ; grad = grad_scale * (pred - target)     (get_loss 'l2', ddpm.py:1196-1205)
__global__ void __launch_bounds__(256)
mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ grad,
                float* loss_sum, long long n, float grad_scale) {
    __shared__ float red[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = pred[i] - target[i];
        acc += d * d;
        grad[i] = grad_scale * d;
    }
    const float t = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(loss_sum, t);
}

// AdamW (torch.optim.AdamW semantics, ddpm.py:1655): decoupled weight decay, bias-corrected moments.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float bc1, float bc2, float grad_scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gr = g[i] * grad_scale;
    float pv = p[i] * (1.0f - lr * weight_decay);
    const float mv = beta1 * m[i] + (1.0f - beta1) * gr;
    const float vv = beta2 * v[i] + (1.0f - beta2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
    pv -= (lr / bc1) * mv / denom;
    p[i] = pv;
}

// Per-segment bookkeeping of mobi_adamw_segments: one thread per segment.
__global__ void adamw_seg_prepare_kernel(const float* __restrict__ flags, int* __restrict__ steps, float* __restrict__ state,
                                         float beta1, float beta2) {
    const int s = threadIdx.x;
    if (s >= 3) return;
    const bool active = flags[s] > 0.f;
    const int st = steps[s] + (active ? 1 : 0);
    steps[s] = st;
    state[3 * s + 0] = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(st)));
    state[3 * s + 1] = static_cast<float>(1.0 - pow(static_cast<double>(beta2), static_cast<double>(st)));
    state[3 * s + 2] = active ? 1.f : 0.f;
}

__global__ void adamw_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, long long b1, long long b2,
                                 const float* __restrict__ state, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float grad_scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = (i >= b1) + (i >= b2);
    if (state[3 * s + 2] == 0.f) return;  // gradient None this step: torch.optim.AdamW skips the parameter
    const float bc1 = state[3 * s], bc2 = state[3 * s + 1];
    const float gr = g[i] * grad_scale;
    float pv = p[i] * (1.0f - lr * weight_decay);
    const float mv = beta1 * m[i] + (1.0f - beta1) * gr;
    const float vv = beta2 * v[i] + (1.0f - beta2) * gr * gr;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
    pv -= (lr / bc1) * mv / denom;
    p[i] = pv;
}

// dst[segment-addressed row r, :] += src[r, :]  (src compact f32 or bf16): joins a gradient computed on the camera-only /
// lidar-only rows (attention.py:246-261) back into the interleaved residual-stream gradient.
__global__ void scatter_add_rows_kernel(const void* __restrict__ src, int is_f32, float* dst, long long rows, int C,
                                        long long seg, long long seg_stride, long long seg_offset) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const long long d = seg > 0 ? (r / seg) * seg_stride + seg_offset + r % seg : r;
    dst[d * C + c] += ld_any(src, is_f32, i);
}

__global__ void latent_input_kernel(mobi_latent_input_args a) {
    const long long total = (long long)a.n * 9 * a.S * a.S;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % a.S);
    long long r = i / a.S;
    const int y = (int)(r % a.S);
    r /= a.S;
    const int c = (int)(r % 9);
    const int img = (int)(r / 9);
    const int ys = y - a.pad, xs = x + a.left;
    float v = 0.f;
    if (ys >= 0 && ys < a.hs && xs >= 0 && xs < a.ws) {
        if (c < 8) {
            const float* mom = c < 4 ? a.moments_gt : a.moments_inpaint;
            const float* nz = c < 4 ? a.noise_gt : a.noise_inpaint;
            const int cc = c & 3;
            const long long plane = (long long)a.hs * a.ws;
            const long long pix = (long long)ys * a.ws + xs;
            const float mean = mom[((long long)img * 8 + cc) * plane + pix];
            float lv = mom[((long long)img * 8 + 4 + cc) * plane + pix];
            lv = fminf(fmaxf(lv, -30.0f), 20.0f);
            const float eps = nz ? nz[((long long)img * 4 + cc) * plane + pix] : 0.f;
            v = a.scale * (mean + expf(0.5f * lv) * eps);
        } else {
            // F.interpolate(mask, size=ws, mode="nearest"): src = floor(dst * in / out), output is ws x ws
            const int my = min((int)floorf(ys * ((float)a.hm / a.ws)), a.hm - 1);
            const int mx = min((int)floorf(xs * ((float)a.wm / a.ws)), a.wm - 1);
            v = a.mask[((long long)img * a.hm + my) * a.wm + mx];
        }
    }
    a.out[((((long long)img * a.row_stride + a.row_offset) * 9 + c) * a.S + y) * a.S + x] = v;
}

// dx = dy * silu'(pre), silu'(a) = s (1 + a (1 - s)), s = sigmoid(a)   (BBoxEmbedder.second_linear, modules.py:195-201)
__global__ void silu_bwd_kernel(const void* __restrict__ pre, int pre_f32, const void* __restrict__ dy, int dy_f32,
                                float* __restrict__ dx, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = ld_any(pre, pre_f32, i);
    const float sg = 1.0f / (1.0f + __expf(-a));
    dx[i] = ld_any(dy, dy_f32, i) * sg * (1.0f + a * (1.0f - sg));
}

__global__ void bbox_renorm_kernel(float* bbox, long long n_points, float W, float left, float S, float pad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    bbox[3 * i] = (bbox[3 * i] * W - left) / S;
    bbox[3 * i + 1] += pad / S;
}

}  // namespace mobi

using namespace mobi;

#define MOBI_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)

extern "C" int mobi_transpose_bf16(const void* in, int32_t in_dtype, void* out, int64_t batch, int32_t rows, int32_t cols,
                                   int64_t ld_in, int64_t in_batch_stride, int64_t ld_out, int64_t out_batch_stride,
                                   void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(in && out && batch > 0 && rows > 0 && cols > 0 && batch < 65536, "mobi_transpose_bf16: bad argument");
    if (in_dtype == MOBI_DTYPE_BF16 && rows % 2 == 0 && cols % 2 == 0 && ld_in % 2 == 0 && in_batch_stride % 2 == 0 &&
        ld_out % 2 == 0 && out_batch_stride % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 3) == 0 && (rows + 63) / 64 < 65536) {
        dim3 g64((cols + 63) / 64, (rows + 63) / 64, (unsigned)batch);
        transpose_bf16x2_kernel<<<g64, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(in),
                                                         reinterpret_cast<uint32_t*>(out), rows, cols, ld_in,
                                                         in_batch_stride, ld_out, out_batch_stride);
        MOBI_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, (unsigned)batch);
    MOBI_CHECK(grid.y < 65536, "mobi_transpose_bf16: rows=%d too large", rows);
    transpose_kernel<<<grid, 256, 0, stream>>>(in, in_dtype == MOBI_DTYPE_F32, reinterpret_cast<__nv_bfloat16*>(out), rows,
                                               cols, ld_in, in_batch_stride, ld_out, out_batch_stride);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_layernorm_bwd(const mobi_layernorm_bwd_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->x && a->dy && a->dx && a->rows > 0, "mobi_layernorm_bwd: bad argument");
    MOBI_CHECK(a->C > 0 && a->C <= 32 * LNB_MAXV, "mobi_layernorm_bwd: C=%d out of range", a->C);
    MOBI_CHECK((a->dgamma == nullptr) == (a->dbeta == nullptr), "mobi_layernorm_bwd: dgamma and dbeta go together");
    const int warps = 8;
    long long blocks = (a->rows + warps - 1) / warps;
    if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
    const size_t smem = a->dgamma ? 2 * a->C * sizeof(float) : 0;
    const bool aligned = a->C % 2 == 0 && reinterpret_cast<uintptr_t>(a->x) % 8 == 0 && reinterpret_cast<uintptr_t>(a->dx) % 8 == 0 &&
                         reinterpret_cast<uintptr_t>(a->dy) % 8 == 0 && (!a->gamma || reinterpret_cast<uintptr_t>(a->gamma) % 8 == 0);
    if (aligned) {
        const bool f32 = a->dy_dtype == MOBI_DTYPE_F32;
#define LNB_LAUNCH(NP, F)                                                                                                \
    ln_bwd2_kernel<NP, F><<<(unsigned)blocks, warps * 32, smem, stream>>>(a->x, a->gamma, a->dy, a->dx, a->dgamma, a->dbeta, \
                                                                         a->rows, a->C, a->seg, a->seg_stride,             \
                                                                         a->seg_offset, a->eps, a->accumulate)
        if (a->C <= 320) { if (f32) LNB_LAUNCH(5, true); else LNB_LAUNCH(5, false); }
        else if (a->C <= 640) { if (f32) LNB_LAUNCH(10, true); else LNB_LAUNCH(10, false); }
        else { if (f32) LNB_LAUNCH(20, true); else LNB_LAUNCH(20, false); }
#undef LNB_LAUNCH
    } else
    ln_bwd_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(a->x, a->gamma, a->dy, a->dy_dtype == MOBI_DTYPE_F32,
                                                                 a->dx, a->dgamma, a->dbeta, a->rows, a->C, a->seg,
                                                                 a->seg_stride, a->seg_offset, a->eps, a->accumulate);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_groupnorm_bwd(const mobi_groupnorm_bwd_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->x1 && a->gamma && a->beta && a->dy && a->dx1, "mobi_groupnorm_bwd: null argument");
    MOBI_CHECK(a->c2 == 0 || (a->x2 && a->dx2), "mobi_groupnorm_bwd: second source / gradient missing");
    MOBI_CHECK(a->groups > 0 && (a->c1 + a->c2) % a->groups == 0, "mobi_groupnorm_bwd: C=%d not divisible by groups=%d",
               a->c1 + a->c2, a->groups);
    dim3 grid(a->groups * GNB_CLUSTER, a->n_img);  // clusters of GNB_CLUSTER CTAs along x
    {
        const int cpg = (a->c1 + a->c2) / a->groups;
        auto al8 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 8 == 0; };
        const bool pairs = cpg % 2 == 0 && cpg / 2 <= 256 && a->c1 % 2 == 0 && a->c2 % 2 == 0 && al8(a->x1) && al8(a->x2) &&
                           al8(a->gamma) && al8(a->beta) && al8(a->dy) && al8(a->dres) && al8(a->dx1) && al8(a->dx2);
        if (pairs) {
            if (a->dy_dtype == MOBI_DTYPE_F32)
                gn_bwd2_kernel<true><<<grid, 256, 0, stream>>>(a->x1, a->x2, a->gamma, a->beta, a->dy, a->dres, a->dx1, a->dx2,
                                                               a->hw, a->c1, a->c2, a->groups, a->silu, a->eps);
            else
                gn_bwd2_kernel<false><<<grid, 256, 0, stream>>>(a->x1, a->x2, a->gamma, a->beta, a->dy, a->dres, a->dx1, a->dx2,
                                                                a->hw, a->c1, a->c2, a->groups, a->silu, a->eps);
            MOBI_CUDA(cudaGetLastError());
            return 0;
        }
    }
    gn_bwd_kernel<<<grid, 256, 0, stream>>>(a->x1, a->x2, a->gamma, a->beta, a->dy, a->dy_dtype == MOBI_DTYPE_F32,
                                            a->dres, a->dx1, a->dx2, a->hw, a->c1, a->c2, a->groups, a->silu, a->eps);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_geglu(const void* g, void* out, int64_t rows, int64_t features, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(g && out && rows > 0 && features > 0, "mobi_geglu: bad argument");
    const long long n = rows * features;
    geglu_fwd_kernel<<<bw_blocks(n, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat162*>(g),
                                                           reinterpret_cast<__nv_bfloat16*>(out), n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_geglu_bwd(const void* g, const void* dh, void* dg, int64_t rows, int64_t features, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(g && dh && dg && rows > 0 && features > 0, "mobi_geglu_bwd: bad argument");
    const long long n = rows * features;
    geglu_bwd_kernel<<<bw_blocks(n, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat162*>(g),
                                                           reinterpret_cast<const __nv_bfloat16*>(dh),
                                                           reinterpret_cast<__nv_bfloat162*>(dg), n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_attn_softmax_bwd(const mobi_attn_softmax_bwd_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->S && a->dP && a->dS && a->dSt && a->Pt && a->stats, "mobi_attn_softmax_bwd: null argument");
    MOBI_CHECK(a->tq > 0 && a->tk > 0 && a->batch > 0 && a->batch < 65536, "mobi_attn_softmax_bwd: bad shape");
    const long long bs = (long long)a->tq * a->tk;
    dim3 g1((a->tq + 7) / 8, a->batch);
    attn_bwd_stats_kernel<<<g1, 256, 0, stream>>>(a->S, a->dP, a->stats, a->tq, a->tk, bs);
    MOBI_CUDA(cudaGetLastError());
    dim3 g2((a->tk + 31) / 32, (a->tq + 31) / 32, a->batch);
    MOBI_CHECK(g2.y < 65536, "mobi_attn_softmax_bwd: tq too large");
    attn_bwd_apply_kernel<<<g2, 256, 0, stream>>>(a->S, a->dP, a->stats, reinterpret_cast<__nv_bfloat16*>(a->dS),
                                                  reinterpret_cast<__nv_bfloat16*>(a->dSt),
                                                  reinterpret_cast<__nv_bfloat16*>(a->Pt), a->tq, a->tk, a->dscale, bs);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_attn_softmax_bwd_lse(const mobi_attn_softmax_bwd_args* a, const float* lse, const void* o,
                                         const void* d_o, int64_t ld, int32_t head_dim, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->S && a->dP && a->dS && a->dSt && a->Pt && a->stats && lse && o && d_o,
               "mobi_attn_softmax_bwd_lse: null argument");
    MOBI_CHECK(a->tq > 0 && a->tk > 0 && a->batch > 0 && a->batch < 65536 && head_dim > 0, "mobi_attn_softmax_bwd_lse: bad shape");
    const long long bs = (long long)a->tq * a->tk;
    dim3 g1((a->tq + 7) / 8, a->batch);
    attn_bwd_stats_from_lse_kernel<<<g1, 256, 0, stream>>>(lse, reinterpret_cast<const __nv_bfloat16*>(o),
                                                           reinterpret_cast<const __nv_bfloat16*>(d_o), a->stats, a->tq, ld,
                                                           head_dim);
    MOBI_CUDA(cudaGetLastError());
    dim3 g2((a->tk + 31) / 32, (a->tq + 31) / 32, a->batch);
    MOBI_CHECK(g2.y < 65536, "mobi_attn_softmax_bwd_lse: tq too large");
    attn_bwd_apply_kernel<<<g2, 256, 0, stream>>>(a->S, a->dP, a->stats, reinterpret_cast<__nv_bfloat16*>(a->dS),
                                                  reinterpret_cast<__nv_bfloat16*>(a->dSt),
                                                  reinterpret_cast<__nv_bfloat16*>(a->Pt), a->tq, a->tk, a->dscale, bs);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_ctx_attn_qspace(const mobi_ctx_attn_qspace_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->q && a->k && a->v, "mobi_ctx_attn_qspace: null argument");
    MOBI_CHECK(a->keys >= 1 && a->keys <= CA_MAXK && a->heads > 0 && a->C % a->heads == 0,
               "mobi_ctx_attn_qspace: keys=%d (1..4), heads=%d, C=%d", a->keys, a->heads, a->C);
    const int D = a->C / a->heads;
    dim3 grid((a->tokens + CA_TOK - 1) / CA_TOK, a->heads, a->batch);
    const size_t smem = (2 * a->keys * D + 2 * a->keys * CA_TOK) * sizeof(float);
    MOBI_CHECK(!a->backward || (a->d_o && a->dq && a->dk && a->dv), "mobi_ctx_attn_qspace: backward needs d_o, dq, dk, dv");
    MOBI_CHECK(a->backward || a->o != nullptr, "mobi_ctx_attn_qspace: forward needs o");
    const bool al16 = a->C % 8 == 0 && reinterpret_cast<uintptr_t>(a->q) % 16 == 0 &&
                      (a->backward ? (reinterpret_cast<uintptr_t>(a->d_o) % 16 == 0 && reinterpret_cast<uintptr_t>(a->dq) % 16 == 0)
                                   : reinterpret_cast<uintptr_t>(a->o) % 16 == 0);
    if (al16 && (D == 40 || D == 80 || D == 160)) {
        const size_t smem_v = smem + (a->backward ? 2 * (size_t)CA_TOK * (D + 8) * sizeof(__nv_bfloat16) : 0);
#define CA_LAUNCH(B, DVv)                                                                                                   \
    do {                                                                                                                    \
        static bool attr_set = false;                                                                                       \
        if (!attr_set) {                                                                                                    \
            MOBI_CUDA(cudaFuncSetAttribute(ctx_attn_qspace_vec_kernel<B, DVv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           100 * 1024));                                                                    \
            attr_set = true;                                                                                                \
        }                                                                                                                   \
        ctx_attn_qspace_vec_kernel<B, DVv><<<grid, CA_TOK, smem_v, stream>>>(                                               \
            reinterpret_cast<const __nv_bfloat16*>(a->q), a->k, a->v, reinterpret_cast<__nv_bfloat16*>(a->o),               \
            reinterpret_cast<const __nv_bfloat16*>(a->d_o), reinterpret_cast<__nv_bfloat16*>(a->dq), a->dk, a->dv,         \
            a->tokens, a->C, a->heads, a->keys);                                                                            \
    } while (0)
        if (a->backward) {
            if (D == 40) CA_LAUNCH(true, 5);
            else if (D == 80) CA_LAUNCH(true, 10);
            else CA_LAUNCH(true, 20);
        } else {
            if (D == 40) CA_LAUNCH(false, 5);
            else if (D == 80) CA_LAUNCH(false, 10);
            else CA_LAUNCH(false, 20);
        }
#undef CA_LAUNCH
        MOBI_CUDA(cudaGetLastError());
        return 0;
    }
    if (a->backward) {
        MOBI_CHECK(a->d_o && a->dq && a->dk && a->dv, "mobi_ctx_attn_qspace: backward needs d_o, dq, dk, dv");
        ctx_attn_qspace_kernel<true><<<grid, CA_TOK, smem, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(a->q), a->k, a->v, nullptr,
            reinterpret_cast<const __nv_bfloat16*>(a->d_o), reinterpret_cast<__nv_bfloat16*>(a->dq), a->dk, a->dv,
            a->tokens, a->C, a->heads, a->keys);
    } else {
        MOBI_CHECK(a->o != nullptr, "mobi_ctx_attn_qspace: forward needs o");
        ctx_attn_qspace_kernel<false><<<grid, CA_TOK, smem, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(a->q), a->k, a->v, reinterpret_cast<__nv_bfloat16*>(a->o), nullptr,
            nullptr, nullptr, nullptr, a->tokens, a->C, a->heads, a->keys);
    }
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_colsum(const void* x, int32_t dtype, int64_t rows, int32_t cols, int64_t ld, int64_t rows_per_group,
                           float* out, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(x && out && rows > 0 && cols > 0, "mobi_colsum: bad argument");
    if (rows_per_group <= 0) rows_per_group = rows;
    MOBI_CHECK(rows % rows_per_group == 0, "mobi_colsum: rows must be a multiple of rows_per_group");
    int chunk = 256;
    while (rows_per_group % chunk != 0) chunk >>= 1;  // a chunk never straddles two groups
    MOBI_CHECK(rows / chunk < 65536, "mobi_colsum: too many row chunks");
    const bool f32 = dtype == MOBI_DTYPE_F32;
    if (cols % 2 == 0 && ld % 2 == 0 && reinterpret_cast<uintptr_t>(x) % (f32 ? 8 : 4) == 0) {
        // enough blocks to fill the machine: halve the chunk while there are fewer than ~2 blocks per SM
        while (chunk > 32 && (long long)((cols + 127) / 128) * (rows / chunk) < 2ll * sm_count()) chunk >>= 1;
        dim3 grid2((cols + 127) / 128, (unsigned)(rows / chunk));
        MOBI_CHECK(rows / chunk < 65536, "mobi_colsum: too many row chunks");
        if (f32) colsum2_kernel<true><<<grid2, 256, 0, stream>>>(x, rows, cols, ld, rows_per_group, chunk, out);
        else colsum2_kernel<false><<<grid2, 256, 0, stream>>>(x, rows, cols, ld, rows_per_group, chunk, out);
        MOBI_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 grid((cols + 31) / 32, (unsigned)(rows / chunk));
    colsum_kernel<<<grid, 256, 0, stream>>>(x, f32, rows, cols, ld, rows_per_group, chunk, out);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_wgrad_small(const float* A, const float* B, float* out, int32_t m, int32_t n, int32_t k, int64_t lda,
                                int64_t ldb, int64_t ldo, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(A && B && out && m > 0 && n > 0 && k > 0, "mobi_wgrad_small: bad argument");
    wgrad_small_kernel<<<bw_blocks((long long)n * k, 256), 256, 0, stream>>>(A, B, out, m, n, k, lda, ldb, ldo);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_zero_insert2x(const void* dy, int32_t in_dtype, void* z, int32_t n, int32_t h, int32_t w, int32_t c,
                                  void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(dy && z && n > 0 && h > 0 && w > 0 && c > 0, "mobi_zero_insert2x: bad argument");
    const long long total = (long long)n * 4 * h * w * c;
    zero_insert2x_kernel<<<bw_blocks(total, 256), 256, 0, stream>>>(dy, in_dtype == MOBI_DTYPE_F32,
                                                                   reinterpret_cast<__nv_bfloat16*>(z), n, h, w, c);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_sum2x2(const float* d, float* out, int32_t n, int32_t h, int32_t w, int32_t c, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(d && out && n > 0 && h > 0 && w > 0 && c > 0, "mobi_sum2x2: bad argument");
    const long long total = (long long)n * h * w * c;
    sum2x2_kernel<<<bw_blocks(total, 256), 256, 0, stream>>>(d, out, n, h, w, c);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac,
                             const int64_t* t, float* out, int32_t batch, int32_t c_total, int32_t c_noised, int32_t hw,
                             void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(x0 && noise && sqrt_ac && sqrt_1mac && t && out && c_noised <= c_total, "mobi_q_sample: bad argument");
    const long long total = (long long)batch * c_total * hw;
    q_sample_kernel<<<bw_blocks(total, 256), 256, 0, stream>>>(x0, noise, sqrt_ac, sqrt_1mac,
                                                              reinterpret_cast<const long long*>(t), out, batch, c_total,
                                                              c_noised, hw);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_mse_grad(const float* pred, const float* target, float* grad, float* loss_sum, int64_t n,
                             float grad_scale, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(pred && target && grad && loss_sum && n > 0, "mobi_mse_grad: bad argument");
    long long blocks = (n + 255) / 256;
    if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
    mse_grad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pred, target, grad, loss_sum, n, grad_scale);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                          float eps, float weight_decay, float bias_corr1, float bias_corr2, float grad_scale,
                          void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(p && g && m && v && n > 0, "mobi_adamw: bad argument");
    adamw_kernel<<<bw_blocks(n, 256), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bias_corr1,
                                                        bias_corr2, grad_scale);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_adamw_segments(float* p, const float* g, float* m, float* v, int64_t n, int64_t bound1, int64_t bound2,
                                   const float* flags, int32_t* steps, float* state, float lr, float beta1, float beta2,
                                   float eps, float weight_decay, float grad_scale, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(p && g && m && v && flags && steps && state && n > 0, "mobi_adamw_segments: bad argument");
    MOBI_CHECK(0 <= bound1 && bound1 <= bound2 && bound2 <= n, "mobi_adamw_segments: segment bounds out of order");
    adamw_seg_prepare_kernel<<<1, 32, 0, stream>>>(flags, steps, state, beta1, beta2);
    MOBI_CUDA(cudaGetLastError());
    adamw_seg_kernel<<<bw_blocks(n, 256), 256, 0, stream>>>(p, g, m, v, n, bound1, bound2, state, lr, beta1, beta2, eps,
                                                           weight_decay, grad_scale);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_scatter_add_rows(const void* src, int32_t src_dtype, float* dst, int64_t rows, int32_t C, int64_t seg,
                                     int64_t seg_stride, int64_t seg_offset, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(src && dst && rows > 0 && C > 0, "mobi_scatter_add_rows: bad argument");
    scatter_add_rows_kernel<<<bw_blocks(rows * C, 256), 256, 0, stream>>>(src, src_dtype == MOBI_DTYPE_F32, dst, rows, C,
                                                                         seg, seg_stride, seg_offset);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_assemble_latent_input(const mobi_latent_input_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->moments_gt && a->moments_inpaint && a->mask && a->out, "mobi_assemble_latent_input: null argument");
    MOBI_CHECK((a->noise_gt == nullptr) == (a->noise_inpaint == nullptr), "mobi_assemble_latent_input: noise_gt and noise_inpaint go together");
    MOBI_CHECK(a->n > 0 && a->hs > 0 && a->ws > 0 && a->hm > 0 && a->wm > 0 && a->S > 0 && a->row_stride >= 1 &&
                   a->row_offset >= 0 && a->row_offset < a->row_stride,
               "mobi_assemble_latent_input: bad shape");
    const long long total = (long long)a->n * 9 * a->S * a->S;
    latent_input_kernel<<<bw_blocks(total, 256), 256, 0, stream>>>(*a);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_bbox_renorm(float* bbox, int64_t n_points, int32_t W, int32_t left, int32_t S, int32_t pad,
                                void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(bbox && n_points > 0 && S > 0, "mobi_bbox_renorm: bad argument");
    bbox_renorm_kernel<<<bw_blocks(n_points, 128), 128, 0, stream>>>(bbox, n_points, (float)W, (float)left, (float)S,
                                                                    (float)pad);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_silu_bwd(const void* pre, int32_t pre_dtype, const void* dy, int32_t dy_dtype, float* dx, int64_t n,
                             void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(pre && dy && dx && n > 0, "mobi_silu_bwd: bad argument");
    silu_bwd_kernel<<<bw_blocks(n, 256), 256, 0, stream>>>(pre, pre_dtype == MOBI_DTYPE_F32, dy, dy_dtype == MOBI_DTYPE_F32,
                                                          dx, n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}
