// HBM-bound kernels of the denoising path: GroupNorm(+SiLU), LayerNorm, timestep embedding, layout
// changes, nearest upsampling, im2col for the strided convolutions, the folded 2-key context attention
// and the fused sampler update.  All are coalesced along the channel (innermost) dimension, vectorised to
// 16-byte accesses where alignment allows, and use warp-shuffle reductions; none stages through shared
// memory unless data is reused.
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace mobi {

// x * sigmoid(x) = x * (0.5 + 0.5 tanh(x / 2)): ONE MUFU op (tanh.approx, <= 2^-11 relative) and three FMA-pipe
// instructions per element instead of MUFU.EX2 + an IEEE division (ncu on gn_stream_kernel: XU pipe 57 %, issue 72 % of a
// kernel that should only stream); the result is rounded to bf16 (2^-9) right after.
__device__ __forceinline__ float silu_f(float x) { return x * fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// ------------------------------------------------------------------------------------------------
// GroupNorm: pass 1 = deterministic per-slab partial sums, pass 2 = finalize + normalise (+SiLU).
// Both passes: warp <-> pixel, lane <-> 4-channel vectors (16-byte loads, 8-byte bf16 stores), GN_K vectors of a
// pixel in flight per lane.
// ------------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_WARPS = GN_THREADS / 32;
constexpr int GN_K = 4;  // float4 vectors per lane per pass: 512 channels per pass

template <bool IN_F32>
__device__ __forceinline__ float4 gn_load4(const void* x1, const void* x2, int c1, int c2, long long pix, int c) {
    // c and c1 are multiples of 4, so a vector never straddles the concatenation boundary
    const void* src = (c < c1) ? x1 : x2;
    const int cc = (c < c1) ? c : c - c1;
    const int cw = (c < c1) ? c1 : c2;
    if (IN_F32) {
        return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + pix * cw + cc);
    } else {
        const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(src) + pix * cw + cc);
        const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
        const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
}

// grid (slabs, n_img). partials layout [n_img][slabs][groups][2] (sum, sumsq)
template <bool IN_F32>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const void* x1, const void* x2, float* partials, int hw, int c1, int c2, int groups,
                int pix_per_slab) {
    extern __shared__ float gn_smem[];  // [GN_WARPS][C][2]
    const int C = c1 + c2;
    const int nv = C >> 2;
    const int cpg = C / groups;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_slab;
    const int p1 = min(hw, p0 + pix_per_slab);
    const long long img_off = (long long)n * hw;

    for (int vbase = 0; vbase < nv; vbase += 32 * GN_K) {
        float4 s[GN_K], q[GN_K];
#pragma unroll
        for (int k = 0; k < GN_K; ++k) s[k] = q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        // two pixels per iteration: 2 * GN_K independent 16-byte loads in flight per lane
        for (int p = p0 + warp; p < p1; p += 2 * GN_WARPS) {
            float4 f[2][GN_K];
            const bool second = (p + GN_WARPS) < p1;
#pragma unroll
            for (int k = 0; k < GN_K; ++k) {
                const int v = vbase + lane + 32 * k;
                f[0][k] = f[1][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v < nv) {
                    f[0][k] = gn_load4<IN_F32>(x1, x2, c1, c2, img_off + p, 4 * v);
                    if (second) f[1][k] = gn_load4<IN_F32>(x1, x2, c1, c2, img_off + p + GN_WARPS, 4 * v);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < GN_K; ++k) {
                    const float4 g = f[u][k];
                    s[k].x += g.x; s[k].y += g.y; s[k].z += g.z; s[k].w += g.w;
                    q[k].x += g.x * g.x; q[k].y += g.y * g.y; q[k].z += g.z * g.z; q[k].w += g.w * g.w;
                }
        }
#pragma unroll
        for (int k = 0; k < GN_K; ++k) {
            const int v = vbase + lane + 32 * k;
            if (v < nv) {
                float4* dst = reinterpret_cast<float4*>(gn_smem + ((size_t)warp * C + 4 * v) * 2);
                dst[0] = make_float4(s[k].x, q[k].x, s[k].y, q[k].y);
                dst[1] = make_float4(s[k].z, q[k].z, s[k].w, q[k].w);
            }
        }
    }
    __syncthreads();
    // 8 threads per group: each sums a strided share of the group's (warp, channel) cells, then a shuffle reduce
    {
        const int g = threadIdx.x >> 3, part = threadIdx.x & 7;
        float s = 0.f, q = 0.f;
        if (g < groups) {
            const int cells = GN_WARPS * cpg;
            for (int i = part; i < cells; i += 8) {
                const int w = i / cpg, c = g * cpg + (i - w * cpg);
                s += gn_smem[((size_t)w * C + c) * 2 + 0];
                q += gn_smem[((size_t)w * C + c) * 2 + 1];
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && part == 0) {
            float* out = partials + (((long long)n * gridDim.x + blockIdx.x) * groups + g) * 2;
            out[0] = s;
            out[1] = q;
        }
    }
}

// grid (slabs2, n_img): each CTA finalises the statistics of its image (cheap) and normalises a slab.
template <bool IN_F32, bool OUT_F32>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const void* x1, const void* x2, const float* __restrict__ gamma, const float* __restrict__ beta,
                void* __restrict__ out_, __nv_bfloat16* __restrict__ out_concat,
                const float* __restrict__ partials, int n_stat_slabs, int hw, int c1, int c2, int groups, float eps,
                int silu, int pix_per_slab) {
    extern __shared__ float gn_smem[];  // scale[C], shift[C]
    __shared__ float s_mean[32], s_rstd[32];
    const int C = c1 + c2;
    const int nv = C >> 2;
    const int cpg = C / groups;
    const int n = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        // 8 threads per group over the stat slabs (groups <= 32)
        const int g = threadIdx.x >> 3, part = threadIdx.x & 7;
        double s = 0.0, q = 0.0;
        if (g < groups) {
            const float* pp = partials + ((long long)n * n_stat_slabs * groups + g) * 2;
            for (int i = part; i < n_stat_slabs; i += 8) {
                s += (double)pp[(long long)i * groups * 2 + 0];
                q += (double)pp[(long long)i * groups * 2 + 1];
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && part == 0) {
            const double cnt = (double)hw * cpg;
            const double mean = s / cnt;
            double var = q / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean[g] = (float)mean;
            s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
    __syncthreads();
    float* s_scale = gn_smem;
    float* s_shift = gn_smem + C;
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        const int g = c / cpg;
        const float a = gamma[c] * s_rstd[g];
        s_scale[c] = a;
        s_shift[c] = beta[c] - s_mean[g] * a;
    }
    __syncthreads();
    const int p0 = blockIdx.x * pix_per_slab;
    const int p1 = min(hw, p0 + pix_per_slab);
    const long long img_off = (long long)n * hw;
    for (int p = p0 + warp; p < p1; p += 2 * GN_WARPS) {
        const bool second = (p + GN_WARPS) < p1;
        for (int vbase = 0; vbase < nv; vbase += 32 * GN_K) {
            float4 f[2][GN_K];
#pragma unroll
            for (int k = 0; k < GN_K; ++k) {
                const int v = vbase + lane + 32 * k;
                if (v < nv) {
                    f[0][k] = gn_load4<IN_F32>(x1, x2, c1, c2, img_off + p, 4 * v);
                    if (second) f[1][k] = gn_load4<IN_F32>(x1, x2, c1, c2, img_off + p + GN_WARPS, 4 * v);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !second) break;
                const long long pp = img_off + p + u * GN_WARPS;
#pragma unroll
                for (int k = 0; k < GN_K; ++k) {
                    const int v = vbase + lane + 32 * k;
                    if (v < nv) {
                        const float4 sc = *reinterpret_cast<const float4*>(s_scale + 4 * v);
                        const float4 sh = *reinterpret_cast<const float4*>(s_shift + 4 * v);
                        const float4 g = f[u][k];
                        float y0 = g.x * sc.x + sh.x, y1 = g.y * sc.y + sh.y;
                        float y2 = g.z * sc.z + sh.z, y3 = g.w * sc.w + sh.w;
                        if (silu) {
                            y0 = silu_f(y0);
                            y1 = silu_f(y1);
                            y2 = silu_f(y2);
                            y3 = silu_f(y3);
                        }
                        const long long o = pp * C + 4 * v;
                        if (OUT_F32)
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + o) = make_float4(y0, y1, y2, y3);
                        else
                            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_) + o) =
                                make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
                        if (out_concat)
                            *reinterpret_cast<uint2*>(out_concat + o) = make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
                    }
                }
            }
        }
    }
}


// Producer-side statistics (mobi_gemm_args.colstats): the GEMM / conv that wrote x1 (and x2) left per-column sums and sums
// of squares for every group of 32 output rows, f32 [2][rows / 32][c].  grid (groups, n_img): one CTA folds the
// (row group, channel) cells of one (image, group) into the sums gn_apply_kernel finalises ([n_img][1][groups][2]), so no
// statistics pass reads the image.  Cells are read with the channel index fastest: consecutive threads, consecutive floats.
__global__ void __launch_bounds__(GN_THREADS)
gn_colstats_fold_kernel(const float* __restrict__ st1, const float* __restrict__ st2, float* __restrict__ partials,
                        int n_img, int hw, int c1, int c2, int groups) {
    __shared__ double red[2][GN_WARPS];
    pdl_wait();
    pdl_trigger();
    const int C = c1 + c2;
    const int cpg = C / groups;
    const int slabs = hw >> 5;
    const int g = blockIdx.x, n = blockIdx.y;
    const long long plane1 = (long long)n_img * slabs * c1, plane2 = (long long)n_img * slabs * c2;
    double s = 0.0, q = 0.0;
    const int cells = slabs * cpg;
    for (int i = threadIdx.x; i < cells; i += GN_THREADS) {
        const int slab = i / cpg;
        const int c = g * cpg + (i - slab * cpg);
        const bool first = c < c1;
        const float* src = first ? st1 : st2;
        const long long idx = ((long long)n * slabs + slab) * (first ? c1 : c2) + (first ? c : c - c1);
        s += (double)src[idx];
        q += (double)src[(first ? plane1 : plane2) + idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = s;
        red[1][warp] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ss = 0.0, qq = 0.0;
#pragma unroll
        for (int w = 0; w < GN_WARPS; ++w) {
            ss += red[0][w];
            qq += red[1][w];
        }
        partials[((long long)n * groups + g) * 2 + 0] = (float)ss;
        partials[((long long)n * groups + g) * 2 + 1] = (float)qq;
    }
}

// Streaming normalise pass for the producer-statistics path: grid (slabs, n_img, sources).  One source (x1 or x2 of the
// concatenation) is read as a flat stream of 16-byte vectors, four independent loads in flight per thread; the per-channel
// scale / shift of this source's channels sit in shared memory.  4 B read + 2 B written per element, nothing else.
constexpr int GNS_UN = 4;

template <bool OUT_F32>
__global__ void __launch_bounds__(GN_THREADS)
gn_stream_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ gamma,
                 const float* __restrict__ beta, void* __restrict__ out_, __nv_bfloat16* __restrict__ out_concat,
                 const float* __restrict__ partials, int hw, int c1, int c2, int groups, float eps, int silu,
                 long long vec_per_cta) {
    extern __shared__ float gn_smem[];  // scale[c_src], shift[c_src]
    __shared__ float s_mean[32], s_rstd[32];
    pdl_wait();
    pdl_trigger();
    const int C = c1 + c2;
    const int cpg = C / groups;
    const int n = blockIdx.y;
    const bool second = blockIdx.z != 0;
    const float* x = second ? x2 : x1;
    const int c_src = second ? c2 : c1, c_off = second ? c1 : 0;
    const int nvs = c_src >> 2;
    if (threadIdx.x < groups) {
        const double cnt = (double)hw * cpg;
        const double sm = (double)partials[((long long)n * groups + threadIdx.x) * 2 + 0];
        const double sq = (double)partials[((long long)n * groups + threadIdx.x) * 2 + 1];
        const double mean = sm / cnt;
        double var = sq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    float* s_scale = gn_smem;
    float* s_shift = gn_smem + c_src;
    for (int c = threadIdx.x; c < c_src; c += GN_THREADS) {
        const int g = (c_off + c) / cpg;
        const float a = gamma[c_off + c] * s_rstd[g];
        s_scale[c] = a;
        s_shift[c] = beta[c_off + c] - s_mean[g] * a;
    }
    __syncthreads();
    const long long total = (long long)hw * nvs;  // vectors of this source in this image
    const long long v0 = (long long)blockIdx.x * vec_per_cta;
    const long long v1 = min(total, v0 + vec_per_cta);
    const float4* src = reinterpret_cast<const float4*>(x) + (long long)n * total;
    for (long long base = v0 + threadIdx.x; base < v1; base += (long long)GN_THREADS * GNS_UN) {
        float4 f[GNS_UN];
#pragma unroll
        for (int k = 0; k < GNS_UN; ++k) {
            const long long idx = base + (long long)k * GN_THREADS;
            if (idx < v1) f[k] = __ldcs(src + idx);   // streamed once: do not keep it in L2 in front of the weights
        }
#pragma unroll
        for (int k = 0; k < GNS_UN; ++k) {
            const long long idx = base + (long long)k * GN_THREADS;
            if (idx >= v1) break;
            const unsigned pixel = (unsigned)(idx / (unsigned)nvs);
            const int v = (int)(idx - (long long)pixel * nvs);
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + 4 * v);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + 4 * v);
            const float4 g = f[k];
            float y0 = g.x * sc.x + sh.x, y1 = g.y * sc.y + sh.y, y2 = g.z * sc.z + sh.z, y3 = g.w * sc.w + sh.w;
            if (silu) {
                y0 = silu_f(y0);
                y1 = silu_f(y1);
                y2 = silu_f(y2);
                y3 = silu_f(y3);
            }
            const long long o = ((long long)n * hw + pixel) * C + c_off + 4 * v;
            if (OUT_F32)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + o) = make_float4(y0, y1, y2, y3);
            else
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_) + o) =
                    make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
            if (out_concat)
                *reinterpret_cast<uint2*>(out_concat + o) = make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Single-pass GroupNorm: one 8-CTA thread-block cluster per (image, chunk of `gpc` groups).  Each CTA owns a pixel slab:
// it accumulates the statistics of its slab, the cluster combines them through distributed shared memory, and the CTA
// normalises the SAME slab again — the second read comes from L2 (the host sizes the chunk so that the slabs of all
// resident clusters fit), so HBM sees 4 B read + 2 B written per element instead of 8 + 2.
// ------------------------------------------------------------------------------------------------
constexpr int GNF_CLUSTER = 8;

// Thread mapping: thread = (pixel lane pl, 4-channel vector v) with v = tid % nvc fixed for the whole kernel, so the
// statistics accumulate in registers, the per-channel scale / shift of phase 2 live in registers, and consecutive
// threads read consecutive 16-byte vectors of a pixel (every lane busy even for a 160-channel chunk).
template <bool IN_F32, bool OUT_F32, int GNF_THREADS, int GNF_UNROLL>
__global__ void __cluster_dims__(GNF_CLUSTER, 1, 1) __launch_bounds__(GNF_THREADS)
gn_fused_kernel(const void* x1, const void* x2, const float* __restrict__ gamma, const float* __restrict__ beta,
                void* __restrict__ out_, __nv_bfloat16* __restrict__ out_concat, int hw, int c1, int c2, int groups,
                int gpc, float eps, int silu) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float part_s[GNF_THREADS * 8];  // [pixel lane][Cc][2], pixel lanes * Cc <= 4 * GNF_THREADS (<= 32 KB)
    __shared__ float slot[32][2];
    __shared__ float s_mean[32], s_rstd[32];
    const int C = c1 + c2;
    const int cpg = C / groups;
    const int Cc = gpc * cpg;
    const int nvc = Cc >> 2;
    const int lanes = GNF_THREADS / nvc;  // pixel lanes
    const int v = threadIdx.x % nvc, pl = threadIdx.x / nvc;
    const bool active = pl < lanes;
    const int part = blockIdx.x % GNF_CLUSTER, chunk = blockIdx.x / GNF_CLUSTER;
    const int cbeg = chunk * gpc * cpg;
    const int ch = cbeg + 4 * v;  // first of this thread's 4 channels
    const int n = blockIdx.y;
    const int pix_per = (hw + GNF_CLUSTER - 1) / GNF_CLUSTER;
    const int p0 = min(hw, part * pix_per), p1 = min(hw, p0 + pix_per);
    const long long img_off = (long long)n * hw;

    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (active) {
        for (int p = p0 + pl; p < p1; p += GNF_UNROLL * lanes) {
            float4 f[GNF_UNROLL];
#pragma unroll
            for (int u = 0; u < GNF_UNROLL; ++u) {
                const int pp = p + u * lanes;
                f[u] = pp < p1 ? gn_load4<IN_F32>(x1, x2, c1, c2, img_off + pp, ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < GNF_UNROLL; ++u) {
                const float4 g = f[u];
                s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
                q.x += g.x * g.x; q.y += g.y * g.y; q.z += g.z * g.z; q.w += g.w * g.w;
            }
        }
        float4* dst = reinterpret_cast<float4*>(part_s + ((size_t)pl * Cc + 4 * v) * 2);
        dst[0] = make_float4(s.x, q.x, s.y, q.y);
        dst[1] = make_float4(s.z, q.z, s.w, q.w);
    }
    __syncthreads();
    {
        const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
        float ss = 0.f, qq = 0.f;
        if (g < gpc) {
            const int cells = lanes * cpg;
            for (int i = sub; i < cells; i += 8) {
                const int w = i / cpg, c = g * cpg + (i - w * cpg);
                ss += part_s[((size_t)w * Cc + c) * 2 + 0];
                qq += part_s[((size_t)w * Cc + c) * 2 + 1];
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            qq += __shfl_xor_sync(0xffffffffu, qq, o);
        }
        if (g < gpc && sub == 0) {
            slot[g][0] = ss;
            slot[g][1] = qq;
        }
    }
    cluster.sync();
    if (threadIdx.x < gpc) {
        double ss = 0.0, qq = 0.0;
        for (int r = 0; r < GNF_CLUSTER; ++r) {
            const float* remote = cluster.map_shared_rank(&slot[0][0], r);
            ss += (double)remote[2 * threadIdx.x];
            qq += (double)remote[2 * threadIdx.x + 1];
        }
        const double cnt = (double)hw * cpg;
        const double mean = ss / cnt;
        double var = qq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    cluster.sync();  // every CTA has read every slot (nobody exits early); s_mean / s_rstd visible in the CTA
    if (!active) return;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g = (4 * v + j) / cpg;
        const float a = gamma[ch + j] * s_rstd[g];
        sc[j] = a;
        sh[j] = beta[ch + j] - s_mean[g] * a;
    }
    for (int p = p0 + pl; p < p1; p += GNF_UNROLL * lanes) {
        float4 f[GNF_UNROLL];
#pragma unroll
        for (int u = 0; u < GNF_UNROLL; ++u) {
            const int pp = p + u * lanes;
            if (pp < p1) f[u] = gn_load4<IN_F32>(x1, x2, c1, c2, img_off + pp, ch);
        }
#pragma unroll
        for (int u = 0; u < GNF_UNROLL; ++u) {
            const int pp = p + u * lanes;
            if (pp >= p1) break;
            const float4 g = f[u];
            float y0 = g.x * sc[0] + sh[0], y1 = g.y * sc[1] + sh[1];
            float y2 = g.z * sc[2] + sh[2], y3 = g.w * sc[3] + sh[3];
            if (silu) {
                y0 = silu_f(y0);
                y1 = silu_f(y1);
                y2 = silu_f(y2);
                y3 = silu_f(y3);
            }
            const long long o = (img_off + pp) * C + ch;
            if (OUT_F32)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_) + o) = make_float4(y0, y1, y2, y3);
            else
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_) + o) =
                    make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
            if (out_concat)
                *reinterpret_cast<uint2*>(out_concat + o) = make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
        }
    }
}

// Groups per cluster for the single-pass kernel, or 0 when the two-kernel path should run: the chunk's channel range
// must be 16-byte aligned and at least 128 contiguous bytes per pixel, and one cluster's slab set ~3 MB so that the
// ~18 clusters resident on 148 SMs re-read their data from L2.
static int gn_fused_gpc(int hw, int C, int groups, int in_bytes, int threads = 512) {
    if (hw < GNF_CLUSTER || groups > 32) return 0;
    const int cpg = C / groups;
    const long long target = 3ll << 20;
    int best = 0;
    for (int gpc = 1; gpc <= groups; gpc <<= 1) {
        if (groups % gpc) continue;
        const long long cc = (long long)gpc * cpg;
        if (cc % 4 || cc * in_bytes < 128) continue;
        if (cc > 4 * threads) break;  // one thread per 4-channel vector
        if (best == 0 || (long long)hw * cc * in_bytes <= target) best = gpc;
    }
    if (best == 0) return 0;
    // a chunk that is far above the target even at the smallest usable size (e.g. VAE images at 512^2) gains nothing
    if ((long long)hw * best * cpg * in_bytes > 4 * target) return 0;
    return best;
}

static int gn_slabs(int hw, int C) {
    // ~128 KB of fp32 input per slab, at least GN_WARPS pixels
    long long pix = (128 * 1024) / ((long long)C * 4);
    if (pix < GN_WARPS) pix = GN_WARPS;
    int slabs = (int)((hw + pix - 1) / pix);
    if (slabs > 256) slabs = 256;
    if (slabs < 1) slabs = 1;
    return slabs;
}

// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, __nv_bfloat16* __restrict__ out, int n,
                                          int dim, float log_max_period) {
    const int half = dim / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * half) return;
    const int r = idx / half, i = idx % half;
    const float freq = expf(-log_max_period * (float)i / (float)half);
    const float arg = (float)t[r] * freq;
    out[(long long)r * dim + i] = __float2bfloat16(cosf(arg));
    out[(long long)r * dim + half + i] = __float2bfloat16(sinf(arg));
    if ((dim & 1) && i == 0) out[(long long)r * dim + dim - 1] = __float2bfloat16(0.f);
}

template <bool IN_F32>
__global__ void silu_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = IN_F32 ? reinterpret_cast<const float*>(x)[i]
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
    out[i] = __float2bfloat16(silu_f(v));
}

// NCHW f32 -> NHWC: tile transpose through shared memory, coalesced on both sides.
template <bool OUT_F32>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, void* __restrict__ out, int c, int hw) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int cc = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (cc < c && p < hw) ? x[((long long)n * c + cc) * hw + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int p = p0 + j, cc = c0 + threadIdx.x;
        if (p < hw && cc < c) {
            const long long o = ((long long)n * hw + p) * c + cc;
            if (OUT_F32)
                reinterpret_cast<float*>(out)[o] = tile[threadIdx.x][j];
            else
                reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16(tile[threadIdx.x][j]);
        }
    }
}

template <bool IN_F32>
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, float* __restrict__ out, int c, int hw) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int p = p0 + j, cc = c0 + threadIdx.x;
        float v = 0.f;
        if (p < hw && cc < c) {
            const long long i = ((long long)n * hw + p) * c + cc;
            v = IN_F32 ? reinterpret_cast<const float*>(x)[i]
                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
        }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int cc = c0 + j, p = p0 + threadIdx.x;
        if (cc < c && p < hw) out[((long long)n * c + cc) * hw + p] = tile[threadIdx.x][j];
    }
}

// nearest x2: each thread produces 2 channels of one output pixel
template <bool IN_F32, bool OUT_F32>
__global__ void upsample2x_kernel(const void* __restrict__ x, void* __restrict__ out, int n, int h, int w, int c) {
    const long long half = c / 2;
    const long long total = (long long)n * (2 * h) * (2 * w) * half;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int v = (int)(i % half);
    long long p = i / half;
    const int ox = (int)(p % (2 * w));
    p /= (2 * w);
    const int oy = (int)(p % (2 * h));
    const int b = (int)(p / (2 * h));
    const long long src = (((long long)b * h + (oy >> 1)) * w + (ox >> 1)) * c + 2 * v;
    const long long dst = (((long long)b * 2 * h + oy) * (2 * w) + ox) * c + 2 * v;
    float2 f;
    if (IN_F32)
        f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(x) + src);
    else
        f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(x) + src));
    if (OUT_F32)
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + dst) = f;
    else
        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(out) + dst) = __floats2bfloat162_rn(f.x, f.y);
}

// im2col: K order (kh, kw, c).  Generic form: one thread per output element.
template <bool IN_F32>
__global__ void im2col_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ out, mobi_im2col_args a) {
    const long long total = (long long)a.n * a.ho * a.wo * a.kpad;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int k = (int)(i % a.kpad);
    long long m = i / a.kpad;
    const int ox = (int)(m % a.wo);
    m /= a.wo;
    const int oy = (int)(m % a.ho);
    const int b = (int)(m / a.ho);
    float v = 0.f;
    if (k < a.kh * a.kw * a.c) {
        const int c = k % a.c;
        const int tap = k / a.c;
        const int kx = tap % a.kw, ky = tap / a.kw;
        const int iy = oy * a.stride + ky - a.pad_top;
        const int ix = ox * a.stride + kx - a.pad_left;
        if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) {
            const long long s = (((long long)b * a.h + iy) * a.w + ix) * a.c + c;
            v = IN_F32 ? reinterpret_cast<const float*>(x)[s]
                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[s]);
        }
    }
    out[i] = __float2bfloat16(v);
}

// Vector form for C % 8 == 0 (the strided Downsample convolutions): one thread per 8 channels of one filter tap of
// one output pixel = one 16-byte store; a warp covers 256 consecutive channels, so loads are coalesced too.
// grid.x = output pixels, threads loop over (tap, channel-octet).
template <bool IN_F32>
__global__ void __launch_bounds__(256)
im2col_vec8_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ out, mobi_im2col_args a) {
    const int m = blockIdx.x;  // output pixel (b, oy, ox)
    const int ox = m % a.wo;
    const int t1 = m / a.wo;
    const int oy = t1 % a.ho;
    const int b = t1 / a.ho;
    const int c8 = a.c >> 3;
    const int units = a.kh * a.kw * c8;
    uint4* orow = reinterpret_cast<uint4*>(out + (long long)m * a.kpad);
    for (int u = threadIdx.x; u < units; u += blockDim.x) {
        const int tap = u / c8, cc = (u - tap * c8) << 3;
        const int ky = tap / a.kw, kx = tap - ky * a.kw;
        const int iy = oy * a.stride + ky - a.pad_top;
        const int ix = ox * a.stride + kx - a.pad_left;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (iy >= 0 && iy < a.h && ix >= 0 && ix < a.w) {
            const long long s = (((long long)b * a.h + iy) * a.w + ix) * a.c + cc;
            if (IN_F32) {
                const float4 f0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + s);
                const float4 f1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + s + 4);
                v = make_uint4(pack_bf16x2(f0.x, f0.y), pack_bf16x2(f0.z, f0.w), pack_bf16x2(f1.x, f1.y),
                               pack_bf16x2(f1.z, f1.w));
            } else {
                v = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + s);
            }
        }
        orow[u] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Folded 2-key context attention: one warp per token.
// ------------------------------------------------------------------------------------------------
constexpr int CTX_MAX_HK = 16;

__global__ void __launch_bounds__(256)
ctx_attention_kernel(const __nv_bfloat16* __restrict__ xn, const float* __restrict__ U, const float* __restrict__ Z,
                     const float* __restrict__ zb, float* __restrict__ x, int tokens, int C, int heads, int keys,
                     long long rows) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / tokens);
    const int hk = heads * keys;
    const float* Ub = U + (long long)b * hk * C;
    const float* Zb = Z + (long long)b * hk * C;
    const __nv_bfloat16* xr = xn + row * C;
    float s[CTX_MAX_HK];
#pragma unroll
    for (int j = 0; j < CTX_MAX_HK; ++j) s[j] = 0.f;
    for (int c = lane * 2; c < C; c += 64) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + c));
#pragma unroll
        for (int j = 0; j < CTX_MAX_HK; ++j) {
            if (j < hk) {
                const float2 u = __ldg(reinterpret_cast<const float2*>(Ub + (long long)j * C + c));
                s[j] += f.x * u.x + f.y * u.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CTX_MAX_HK; ++j) s[j] = warp_sum(s[j]);
    // softmax over the `keys` scores of each head (layout j = key * heads + head)
    float pr[CTX_MAX_HK];
#pragma unroll
    for (int h = 0; h < CTX_MAX_HK; ++h) {
        if (h < heads) {
            float mx = -INFINITY;
            for (int k = 0; k < keys; ++k) mx = fmaxf(mx, s[k * heads + h]);
            float den = 0.f;
            for (int k = 0; k < keys; ++k) {
                const float e = __expf(s[k * heads + h] - mx);
                pr[k * heads + h] = e;
                den += e;
            }
            const float inv = 1.0f / den;
            for (int k = 0; k < keys; ++k) pr[k * heads + h] *= inv;
        }
    }
    float* xo = x + row * C;
    for (int c = lane * 2; c < C; c += 64) {
        float2 acc = *reinterpret_cast<const float2*>(zb + c);
#pragma unroll
        for (int j = 0; j < CTX_MAX_HK; ++j) {
            if (j < hk) {
                const float2 z = __ldg(reinterpret_cast<const float2*>(Zb + (long long)j * C + c));
                acc.x += pr[j] * z.x;
                acc.y += pr[j] * z.y;
            }
        }
        float2 cur = *reinterpret_cast<float2*>(xo + c);
        cur.x += acc.x;
        cur.y += acc.y;
        *reinterpret_cast<float2*>(xo + c) = cur;
    }
}

// ------------------------------------------------------------------------------------------------
// Sampler kernels (NCHW f32, tiny tensors: launch-bound, so everything is one kernel)
// ------------------------------------------------------------------------------------------------
__global__ void sampler_update_kernel(mobi_sampler_args a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float e;
    if (a.cfg) {
        const float eu = a.eps[i];
        const float ec = a.eps[a.n + i];
        e = eu + a.scale * (ec - eu);
    } else {
        e = a.eps[i];
    }
    if (a.e_out) a.e_out[i] = e;
    float ep = a.c0 * e;
    if (a.old1) ep += a.c1 * a.old1[i];
    if (a.old2) ep += a.c2 * a.old2[i];
    if (a.old3) ep += a.c3 * a.old3[i];
    const float xv = a.x[i];
    const float pred = (xv - a.sqrt_one_minus_at * ep) / a.sqrt_at;
    float xp = a.sqrt_a_prev * pred + a.dir_coef * ep;
    if (a.noise) xp += a.sigma_temp * a.noise[i];
    a.pred_x0[i] = pred;
    a.x_prev[i] = xp;
}

__global__ void assemble_input_kernel(mobi_assemble_args a) {
    // one thread per (b, pixel)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.B * a.hw;
    if (i >= total) return;
    const int b = (int)(i / a.hw);
    const int p = (int)(i % a.hw);
    const int ctot = 4 + a.rest_c;
    float xv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const long long xi = ((long long)b * 4 + c) * a.hw + p;
        float v = a.x[xi];
        if (a.blend_mask) {
            const float m = a.blend_mask[((long long)b * a.blend_c + (a.blend_c == 1 ? 0 : c)) * a.hw + p];
            const float xo = a.sqrt_ac * a.blend_x0[xi] + a.sqrt_1mac * a.blend_noise[xi];
            v = xo * m + (1.0f - m) * v;
            a.x[xi] = v;
        }
        xv[c] = v;
    }
    const int reps = a.cfg ? 2 : 1;
    for (int r = 0; r < reps; ++r) {
        float* dst = a.x_in + ((long long)(r * a.B + b) * ctot) * a.hw + p;
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[(long long)c * a.hw] = xv[c];
        if (a.inpaint_mask != nullptr) {  // test_model_kwargs: 4 inpaint_image channels + 1 mask channel
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[(long long)(4 + c) * a.hw] = a.inpaint_image[((long long)b * 4 + c) * a.hw + p];
            dst[(long long)8 * a.hw] = a.inpaint_mask[(long long)b * a.hw + p];
        } else {
            // generic `rest` tensor [B, rest_c, H, W] passed through inpaint_image
            for (int c = 0; c < a.rest_c; ++c)
                dst[(long long)(4 + c) * a.hw] = a.inpaint_image[((long long)b * a.rest_c + c) * a.hw + p];
        }
    }
}

__global__ void add_f32_kernel(const float* a, const float* b, float* o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}
__global__ void scale_f32_kernel(const float* a, float s, float* o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] * s;
}
__global__ void cast_bf16_kernel(const float* a, __nv_bfloat16* o, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = __float2bfloat16(a[i]);
}

// Row softmax of a score matrix that was produced by a GEMM (VAE AttnBlock, model.py:184-195: one head, d = C = 512,
// too wide for the fused kernels' TMEM budget).  Scores are already scaled by C^-0.5 * log2(e) (folded into the q
// projection), so p = 2^(s - max) / sum.  One CTA per row; the row is read three times (max, sum, write) out of L2.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int cols, long long ld_s, long long ld_p) {
    __shared__ float red[8];
    __shared__ float bcast;
    const float* row = s + (long long)blockIdx.x * ld_s;
    __nv_bfloat16* out = p + (long long)blockIdx.x * ld_p;
    const int nv = cols >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nv; i += 256) {
        const float4 v = reinterpret_cast<const float4*>(row)[i];
        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        bcast = m;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int i = threadIdx.x; i < nv; i += 256) {
        const float4 v = reinterpret_cast<const float4*>(row)[i];
        sum += (exp2f(v.x - mx) + exp2f(v.y - mx)) + (exp2f(v.z - mx) + exp2f(v.w - mx));
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bcast = 1.0f / t;
    }
    __syncthreads();
    const float inv = bcast;
    for (int i = threadIdx.x; i < nv; i += 256) {
        const float4 v = reinterpret_cast<const float4*>(row)[i];
        reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(exp2f(v.x - mx) * inv, exp2f(v.y - mx) * inv),
                                                      pack_bf16x2(exp2f(v.z - mx) * inv, exp2f(v.w - mx) * inv));
    }
}

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace mobi

using namespace mobi;

extern "C" int64_t mobi_groupnorm_scratch_bytes(int32_t n_img, int32_t hw, int32_t c, int32_t groups) {
    return (int64_t)n_img * gn_slabs(hw, c) * groups * 2 * sizeof(float);
}

extern "C" int32_t mobi_groupnorm_launches(int32_t hw, int32_t c, int32_t groups, int32_t in_dtype, int32_t force_two_pass) {
    if (groups <= 0 || c % groups) return 2;
    return (!force_two_pass && gn_fused_gpc(hw, c, groups, in_dtype == MOBI_DTYPE_F32 ? 4 : 2) > 0) ? 1 : 2;
}

extern "C" int mobi_groupnorm(const mobi_groupnorm_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->x1 && a->gamma && a->beta && a->out && a->partials, "mobi_groupnorm: null argument");
    const int C = a->c1 + a->c2;
    MOBI_CHECK(a->groups > 0 && a->groups <= 32 && C % a->groups == 0, "mobi_groupnorm: C=%d groups=%d", C, a->groups);
    MOBI_CHECK(a->c1 % 4 == 0 && a->c2 % 4 == 0,
               "mobi_groupnorm: channel counts of both inputs must be multiples of 4 (c1=%d c2=%d)", a->c1, a->c2);
    MOBI_CHECK(a->c2 == 0 || a->x2 != nullptr, "mobi_groupnorm: c2 > 0 needs x2");
    const bool f32 = a->in_dtype == MOBI_DTYPE_F32;
    const bool of32 = a->out_dtype == MOBI_DTYPE_F32;
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(gn_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        MOBI_CUDA(cudaFuncSetAttribute(gn_stats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
    }
    // apply pass: finer slabs so small images still fill the machine
    long long pix2 = (128 * 1024) / ((long long)C * 4);
    if (pix2 < GN_WARPS) pix2 = GN_WARPS;
    const int slabs2 = (int)((a->hw + pix2 - 1) / pix2);
    const int pix_per_slab2 = (a->hw + slabs2 - 1) / slabs2;
    const dim3 grid2(slabs2, a->n_img);
    const size_t smem2 = (size_t)2 * C * sizeof(float);
#define GN_APPLY(INF, OUTF, NSLABS)                                                                                    \
    gn_apply_kernel<INF, OUTF><<<grid2, GN_THREADS, smem2, stream>>>(                                                  \
        a->x1, a->x2, a->gamma, a->beta, a->out, reinterpret_cast<__nv_bfloat16*>(a->out_concat), a->partials, NSLABS, \
        a->hw, a->c1, a->c2, a->groups, a->eps, a->silu, pix_per_slab2)
#define GN_APPLY_ANY(NSLABS)                       \
    do {                                           \
        if (f32 && of32) GN_APPLY(true, true, NSLABS);     \
        else if (f32) GN_APPLY(true, false, NSLABS);       \
        else if (of32) GN_APPLY(false, true, NSLABS);      \
        else GN_APPLY(false, false, NSLABS);               \
    } while (0)
    if (a->colstats1 != nullptr) {
        // statistics came with the producer's epilogue: fold the column sums, then ONE streaming pass over the image
        MOBI_CHECK(a->hw % 32 == 0 && (a->c2 == 0 || a->colstats2 != nullptr),
                   "mobi_groupnorm: colstats need hw %% 32 == 0 and statistics for both inputs");
        MOBI_CUDA(launch_pdl(gn_colstats_fold_kernel, dim3(a->groups, a->n_img), dim3(GN_THREADS), (size_t)0, stream,
                             a->colstats1, a->colstats2, a->partials, a->n_img, a->hw, a->c1, a->c2, a->groups));
        if (f32) {
            const int cmax = a->c1 > a->c2 ? a->c1 : a->c2;
            const long long vec_img = (long long)a->hw * (cmax >> 2);
            // >= 4 rounds of GN_THREADS x GNS_UN vectors per CTA, at most ~16 CTAs per SM over the grid
            long long per = (long long)GN_THREADS * GNS_UN * 4;
            const long long want = (148ll * 16 + a->n_img - 1) / a->n_img;
            if ((vec_img + per - 1) / per > want) per = ((vec_img + want - 1) / want + GN_THREADS * GNS_UN - 1) / (GN_THREADS * GNS_UN) * (GN_THREADS * GNS_UN);
            const dim3 grid((unsigned)((vec_img + per - 1) / per), a->n_img, a->c2 > 0 ? 2 : 1);
            const size_t smem = (size_t)2 * cmax * sizeof(float);
            if (of32)
                MOBI_CUDA(launch_pdl(gn_stream_kernel<true>, grid, dim3(GN_THREADS), smem, stream,
                                     reinterpret_cast<const float*>(a->x1), reinterpret_cast<const float*>(a->x2), a->gamma,
                                     a->beta, a->out, reinterpret_cast<__nv_bfloat16*>(a->out_concat), a->partials, a->hw,
                                     a->c1, a->c2, a->groups, a->eps, a->silu, per));
            else
                MOBI_CUDA(launch_pdl(gn_stream_kernel<false>, grid, dim3(GN_THREADS), smem, stream,
                                     reinterpret_cast<const float*>(a->x1), reinterpret_cast<const float*>(a->x2), a->gamma,
                                     a->beta, a->out, reinterpret_cast<__nv_bfloat16*>(a->out_concat), a->partials, a->hw,
                                     a->c1, a->c2, a->groups, a->eps, a->silu, per));
        } else {
            GN_APPLY_ANY(1);
        }
        MOBI_CUDA(cudaGetLastError());
        return 0;
    }
    const int gpc = a->force_two_pass ? 0 : gn_fused_gpc(a->hw, C, a->groups, f32 ? 4 : 2);
    if (gpc > 0) {
        dim3 grid(GNF_CLUSTER * (a->groups / gpc), a->n_img);
#define GN_FUSED_V(INF, OUTF, TH, UN)                                                                              \
    gn_fused_kernel<INF, OUTF, TH, UN><<<grid, TH, 0, stream>>>(a->x1, a->x2, a->gamma, a->beta, a->out,           \
                                                                reinterpret_cast<__nv_bfloat16*>(a->out_concat),    \
                                                                a->hw, a->c1, a->c2, a->groups, gpc, a->eps, a->silu)
    // 512 threads x 4 pixels in flight measured best of (256, 8), (512, 4), (512, 8), (1024, 2) at every UNet shape
#define GN_FUSED(INF, OUTF) GN_FUSED_V(INF, OUTF, 512, 4)
        if (f32 && of32) GN_FUSED(true, true);
        else if (f32) GN_FUSED(true, false);
        else if (of32) GN_FUSED(false, true);
        else GN_FUSED(false, false);
#undef GN_FUSED
#undef GN_FUSED_V
        MOBI_CUDA(cudaGetLastError());
        return 0;
    }
    const int slabs = gn_slabs(a->hw, C);
    const int pix_per_slab = (a->hw + slabs - 1) / slabs;
    const size_t smem1 = (size_t)GN_WARPS * C * 2 * sizeof(float);
    MOBI_CHECK(smem1 <= 160 * 1024, "mobi_groupnorm: C=%d too large", C);
    dim3 grid1(slabs, a->n_img);
    if (f32)
        gn_stats_kernel<true><<<grid1, GN_THREADS, smem1, stream>>>(a->x1, a->x2, a->partials, a->hw, a->c1, a->c2,
                                                                    a->groups, pix_per_slab);
    else
        gn_stats_kernel<false><<<grid1, GN_THREADS, smem1, stream>>>(a->x1, a->x2, a->partials, a->hw, a->c1, a->c2,
                                                                     a->groups, pix_per_slab);
    MOBI_CUDA(cudaGetLastError());
    GN_APPLY_ANY(slabs);
#undef GN_APPLY_ANY
#undef GN_APPLY
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_timestep_embedding(const int64_t* t, void* out, int32_t n, int32_t dim, float max_period,
                                       void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(t && out && n > 0 && dim >= 2, "mobi_timestep_embedding: bad argument");
    const long long total = (long long)n * (dim / 2);
    timestep_embedding_kernel<<<blocks_for(total, 128), 128, 0, stream>>>(t, reinterpret_cast<__nv_bfloat16*>(out), n,
                                                                          dim, logf(max_period));
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

namespace mobi {
__global__ void fourier_embed_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int dims,
                                     int num_freqs, long long ld_out) {
    const int per_row = dims * (1 + 2 * num_freqs);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld_out) return;
    const long long r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    float v = 0.f;
    if (c < per_row) {
        const int blk = c / dims, d = c - blk * dims;   // block 0 = identity, then (sin, cos) per frequency
        const float xv = x[r * dims + d];
        if (blk == 0) v = xv;
        else {
            const float f = exp2f((float)((blk - 1) >> 1));
            v = ((blk - 1) & 1) ? cosf(xv * f) : sinf(xv * f);
        }
    }
    out[i] = __float2bfloat16(v);
}
}  // namespace mobi

extern "C" int mobi_fourier_embed(const float* x, void* out, int64_t rows, int32_t dims, int32_t num_freqs, int64_t ld_out,
                                  void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && out && rows > 0 && dims > 0 && num_freqs >= 0 && ld_out >= dims * (1 + 2 * num_freqs),
               "mobi_fourier_embed: bad argument");
    fourier_embed_kernel<<<blocks_for(rows * ld_out, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), rows,
                                                                            dims, num_freqs, ld_out);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_silu(const void* x, int32_t in_dtype, void* out, int64_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && out, "mobi_silu: null argument");
    if (n == 0) return 0;
    if (in_dtype == MOBI_DTYPE_F32)
        silu_kernel<true><<<blocks_for(n, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), n);
    else
        silu_kernel<false><<<blocks_for(n, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_nchw_to_nhwc(const float* x, void* out, int32_t out_dtype, int32_t n, int32_t c, int32_t hw,
                                 void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && out && n > 0 && c > 0 && hw > 0, "mobi_nchw_to_nhwc: bad argument");
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    if (out_dtype == MOBI_DTYPE_F32)
        nchw_to_nhwc_kernel<true><<<grid, block, 0, stream>>>(x, out, c, hw);
    else
        nchw_to_nhwc_kernel<false><<<grid, block, 0, stream>>>(x, out, c, hw);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_nhwc_to_nchw(const void* x, int32_t in_dtype, float* out, int32_t n, int32_t c, int32_t hw,
                                 void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && out && n > 0 && c > 0 && hw > 0, "mobi_nhwc_to_nchw: bad argument");
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    if (in_dtype == MOBI_DTYPE_F32)
        nhwc_to_nchw_kernel<true><<<grid, block, 0, stream>>>(x, out, c, hw);
    else
        nhwc_to_nchw_kernel<false><<<grid, block, 0, stream>>>(x, out, c, hw);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_upsample_nearest2x(const void* x, int32_t in_dtype, void* out, int32_t out_dtype, int32_t n,
                                       int32_t h, int32_t w, int32_t c, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && out && c % 2 == 0, "mobi_upsample_nearest2x: bad argument (C must be even)");
    const long long total = (long long)n * 4 * h * w * (c / 2);
    const unsigned blocks = blocks_for(total, 256);
    const bool i32 = in_dtype == MOBI_DTYPE_F32, o32 = out_dtype == MOBI_DTYPE_F32;
    if (i32 && o32)
        upsample2x_kernel<true, true><<<blocks, 256, 0, stream>>>(x, out, n, h, w, c);
    else if (i32 && !o32)
        upsample2x_kernel<true, false><<<blocks, 256, 0, stream>>>(x, out, n, h, w, c);
    else if (!i32 && o32)
        upsample2x_kernel<false, true><<<blocks, 256, 0, stream>>>(x, out, n, h, w, c);
    else
        upsample2x_kernel<false, false><<<blocks, 256, 0, stream>>>(x, out, n, h, w, c);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_im2col(const mobi_im2col_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->x && a->out, "mobi_im2col: null argument");
    MOBI_CHECK(a->kpad >= a->kh * a->kw * a->c && a->kpad % 8 == 0, "mobi_im2col: kpad=%d too small or not %%8", a->kpad);
    const long long total = (long long)a->n * a->ho * a->wo * a->kpad;
    if (a->c % 8 == 0 && a->kpad == a->kh * a->kw * a->c && (reinterpret_cast<uintptr_t>(a->x) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(a->out) & 15) == 0) {
        const unsigned pixels = (unsigned)((long long)a->n * a->ho * a->wo);
        if (a->in_dtype == MOBI_DTYPE_F32)
            im2col_vec8_kernel<true><<<pixels, 256, 0, stream>>>(a->x, reinterpret_cast<__nv_bfloat16*>(a->out), *a);
        else
            im2col_vec8_kernel<false><<<pixels, 256, 0, stream>>>(a->x, reinterpret_cast<__nv_bfloat16*>(a->out), *a);
        MOBI_CUDA(cudaGetLastError());
        return 0;
    }
    if (a->in_dtype == MOBI_DTYPE_F32)
        im2col_kernel<true><<<blocks_for(total, 256), 256, 0, stream>>>(a->x, reinterpret_cast<__nv_bfloat16*>(a->out), *a);
    else
        im2col_kernel<false><<<blocks_for(total, 256), 256, 0, stream>>>(a->x, reinterpret_cast<__nv_bfloat16*>(a->out), *a);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_ctx_attention(const mobi_ctx_attn_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->xn && a->U && a->Z && a->zb && a->x, "mobi_ctx_attention: null argument");
    MOBI_CHECK(a->heads * a->keys <= CTX_MAX_HK && a->C % 2 == 0, "mobi_ctx_attention: heads*keys=%d > %d",
               a->heads * a->keys, CTX_MAX_HK);
    const long long rows = (long long)a->batch * a->tokens;
    const int warps = 8;
    ctx_attention_kernel<<<blocks_for(rows, warps), warps * 32, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->xn), a->U, a->Z, a->zb, a->x, a->tokens, a->C, a->heads, a->keys, rows);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_softmax_rows(const float* s, void* p, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p,
                                 void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(s && p && rows > 0 && cols > 0, "mobi_softmax_rows: bad argument");
    MOBI_CHECK(cols % 4 == 0 && ld_s % 4 == 0 && ld_p % 4 == 0 && rows < (1ll << 31),
               "mobi_softmax_rows: cols=%d and row strides must be multiples of 4", cols);
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(s, reinterpret_cast<__nv_bfloat16*>(p), cols, ld_s, ld_p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_sampler_update(const mobi_sampler_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->eps && a->x && a->x_prev && a->pred_x0 && a->n > 0, "mobi_sampler_update: bad argument");
    sampler_update_kernel<<<blocks_for(a->n, 256), 256, 0, stream>>>(*a);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_assemble_input(const mobi_assemble_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->x && a->x_in && a->inpaint_image, "mobi_assemble_input: null argument");
    MOBI_CHECK(a->inpaint_mask == nullptr || a->rest_c == 5, "mobi_assemble_input: image + mask means rest_c = 5 (got %d)",
               a->rest_c);
    MOBI_CHECK(!a->blend_mask || (a->blend_x0 && a->blend_noise), "mobi_assemble_input: blend needs x0 and noise");
    const long long total = (long long)a->B * a->hw;
    assemble_input_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(*a);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    add_f32_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(a, b, out, n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}
extern "C" int mobi_scale_f32(const float* x, float s, float* out, int64_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    scale_f32_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(x, s, out, n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}
namespace mobi {
// dst[i] = bf16(src[i] * scale of the segment that holds i): segment j covers [seg_start[j], seg_start[j + 1]) (the last one
// runs to n); starts are sorted and multiples of 4.  One thread = one 16-byte load; a block first checks whether its 1024
// elements lie in a single segment (nearly always: segments are whole weight matrices).
__global__ void __launch_bounds__(256)
cast_bf16_segments_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n,
                          const long long* __restrict__ seg_start, const float* __restrict__ seg_scale, int nseg) {
    auto find = [&](long long i) {  // last segment whose start is <= i
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_start[mid] <= i) lo = mid;
            else hi = mid - 1;
        }
        return lo;
    };
    __shared__ int s_seg;
    __shared__ int s_uniform;
    const long long b0 = (long long)blockIdx.x * 1024;
    if (threadIdx.x == 0) {
        const int j = find(b0);
        s_seg = j;
        s_uniform = (j + 1 >= nseg) || (seg_start[j + 1] >= min(n, b0 + 1024));
    }
    __syncthreads();
    const long long i = b0 + 4 * threadIdx.x;
    if (i >= n) return;
    const float sc = seg_scale[s_uniform ? s_seg : find(i)];
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16x2(v.x * sc, v.y * sc), pack_bf16x2(v.z * sc, v.w * sc));
}
}  // namespace mobi

namespace mobi {
// x[i] *= scale of the segment that holds i, touching only the blocks whose segment scale differs from 1 (same segment
// table as the cast above): turns the gradients w.r.t. the scaled query projections into gradients w.r.t. to_q.weight.
__global__ void __launch_bounds__(256)
scale_segments_kernel(float* __restrict__ x, long long n, const long long* __restrict__ seg_start,
                      const float* __restrict__ seg_scale, int nseg) {
    auto find = [&](long long i) {
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_start[mid] <= i) lo = mid;
            else hi = mid - 1;
        }
        return lo;
    };
    __shared__ int s_seg;
    __shared__ int s_uniform;
    const long long b0 = (long long)blockIdx.x * 1024;
    if (threadIdx.x == 0) {
        const int j = find(b0);
        s_seg = j;
        s_uniform = (j + 1 >= nseg) || (seg_start[j + 1] >= min(n, b0 + 1024));
    }
    __syncthreads();
    if (s_uniform && seg_scale[s_seg] == 1.0f) return;
    const long long i = b0 + 4 * threadIdx.x;
    if (i >= n) return;
    const float sc = seg_scale[s_uniform ? s_seg : find(i)];
    if (sc == 1.0f) return;
    float4 v = *reinterpret_cast<float4*>(x + i);
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    *reinterpret_cast<float4*>(x + i) = v;
}
}  // namespace mobi

extern "C" int mobi_scale_segments(float* x, int64_t n, const int64_t* seg_start, const float* seg_scale, int32_t nseg,
                                   void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && seg_start && seg_scale && nseg > 0 && n > 0 && n % 4 == 0, "mobi_scale_segments: bad argument");
    MOBI_CHECK(reinterpret_cast<uintptr_t>(x) % 16 == 0, "mobi_scale_segments: buffer must be 16-byte aligned");
    mobi::scale_segments_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(
        x, n, reinterpret_cast<const long long*>(seg_start), seg_scale, nseg);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_cast_bf16_segments(const float* src, void* dst, int64_t n, const int64_t* seg_start, const float* seg_scale,
                                       int32_t nseg, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(src && dst && seg_start && seg_scale && nseg > 0 && n > 0 && n % 4 == 0, "mobi_cast_bf16_segments: bad argument");
    MOBI_CHECK(reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 8 == 0,
               "mobi_cast_bf16_segments: buffers must be 16-byte (src) / 8-byte (dst) aligned");
    mobi::cast_bf16_segments_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(
        src, reinterpret_cast<__nv_bfloat16*>(dst), n, reinterpret_cast<const long long*>(seg_start), seg_scale, nseg);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_cast_bf16(const float* x, void* out, int64_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (n == 0) return 0;
    cast_bf16_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), n);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}
