// Range-view post-processing that follows the lidar decode: the reference does this on the host
// in NumPy / cv2 / numba, one sample at a time (scripts/inference_test_bench.py:567-629); here it stays in HBM.
//   range_map_kernel        depth_normalization / inverse_depth_normalization (ldm/data/utils.py:537-580) and the
//                           intensity un-normalisation of ddpm.py:1541, optionally after the clamp of ddpm.py:1504
//   range_undo_kernel       LidarConverter.undo_default_transforms, batched as postprocess_range_depth_int does
//                           (lidar_converter.py:436-485, 230-287; utils.py:471-505): shrink the square crop to
//                           [H, width_crop] (avg pool when divisible, else cv2's INTER_NEAREST rule) and paste it into the
//                           original sweep at crop_left % W with wrap-around; the maps above can be fused into the load
//   range_cloud_kernel      LidarConverter.range2pcd (lidar_converter.py:122-176) with ORDERED compaction, and — in
//                           composite mode — the whole save_samples sequence: point-in-box instance mask of the generated
//                           object (box_np_ops.py:406-427, 712-771), paste into the original sweep, edited cloud
//                           [N, 5] = (x, y, z, intensity, beam index).  One 8-CTA cluster per sweep; the per-CTA point
//                           counts meet through distributed shared memory so the cloud keeps the reference's pixel order.
//   points_in_boxes_kernel  points_in_bbox_corners (box_np_ops.py:453-471) for [N, 3+] points x [M, 8, 3] corners
// Every decision (piece selection, validity, inside test) is taken on fp32 values computed in the reference's operation
// order with explicitly rounded intrinsics (no FMA contraction), so the integer results are reproducible bit for bit;
// sin / cos are the full-precision CUDA ones (<= 2 ulp), which is where coordinates may differ in the last bits.
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include <algorithm>
#include <cooperative_groups.h>

namespace mobi {

// ---- scalar maps (utils.py:537-580, ddpm.py:1541), fp32, reference operation order -----------------------------------
__device__ __forceinline__ float depth_norm_fwd(float d, float mn, float mx, float a, float two_a, float one_m_a) {
    if (d >= mn && d <= mx) return __fadd_rn(-a, __fdiv_rn(__fmul_rn(two_a, __fsub_rn(d, mn)), __fsub_rn(mx, mn)));
    if (d >= -1.f && d < mn) return __fadd_rn(-1.f, __fdiv_rn(__fmul_rn(one_m_a, __fadd_rn(d, 1.f)), __fadd_rn(mn, 1.f)));
    if (d > mx && d <= 1.f) return __fadd_rn(a, __fdiv_rn(__fmul_rn(one_m_a, __fsub_rn(d, mx)), __fsub_rn(1.f, mx)));
    return d;  // outside [-1, 1] (or NaN): the reference leaves torch.empty_like garbage, we pass the value through
}

__device__ __forceinline__ float depth_norm_inv(float x, float mn, float mx, float a, float two_a, float a_m_1, float one_m_a) {
    if (x >= -a && x <= a) return __fadd_rn(mn, __fdiv_rn(__fmul_rn(__fadd_rn(x, a), __fsub_rn(mx, mn)), two_a));
    if (x >= -1.f && x < -a) return __fadd_rn(-1.f, __fdiv_rn(__fmul_rn(-__fadd_rn(x, 1.f), __fadd_rn(mn, 1.f)), a_m_1));
    if (x > a && x <= 1.f) return __fadd_rn(mx, __fdiv_rn(__fmul_rn(__fsub_rn(x, a), __fsub_rn(1.f, mx)), one_m_a));
    return x;
}

__device__ __forceinline__ float int_unnorm(float x) {
    const float y = __fsub_rn(__fmul_rn(-0.5f, logf(__fsub_rn(1.f, __fmul_rn(__fadd_rn(x, 1.f), 0.5f)))), 1.f);
    return fminf(fmaxf(y, -1.f), 1.f);
}

struct MapConst {
    float a, two_a, a_m_1, one_m_a;
};

static MapConst map_const(double a) {  // the reference forms 2 * alpha, alpha - 1, 1 - alpha in Python doubles
    return MapConst{(float)a, (float)(2.0 * a), (float)(a - 1.0), (float)(1.0 - a)};
}

__device__ __forceinline__ float apply_map(float v, int mode, int clamp, float mn, float mx, const MapConst& k) {
    if (clamp) v = fminf(fmaxf(v, -1.f), 1.f);
    if (mode == MOBI_RANGE_MAP_DEPTH_NORM) return depth_norm_fwd(v, mn, mx, k.a, k.two_a, k.one_m_a);
    if (mode == MOBI_RANGE_MAP_DEPTH_UNNORM) return depth_norm_inv(v, mn, mx, k.a, k.two_a, k.a_m_1, k.one_m_a);
    if (mode == MOBI_RANGE_MAP_INT_UNNORM) return int_unnorm(v);
    return v;
}

__global__ void __launch_bounds__(256)
range_map_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ min_d,
                 const float* __restrict__ max_d, long long n, long long in_stride, long long out_stride, int mode,
                 int clamp, MapConst k) {
    const int b = blockIdx.y;
    const float mn = min_d ? min_d[b] : 0.f, mx = max_d ? max_d[b] : 0.f;
    const float* src = in + b * in_stride;
    float* dst = out + b * out_stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = apply_map(src[i], mode, clamp, mn, mx, k);
}

// ---- window sums of the avg-pool shrink ---------------------------------------------------------------------------------
// The per-pixel map of the pooled path with everything that is constant per sample hoisted: for the inverse depth map the
// three pieces share ONE shape, base + (+-(x + off)) * mul / div, so a pixel costs a few selects and a single division.
struct PixelMap {
    int mode, clamp;
    float mn, mx, a, two_a, a_m_1, one_m_a, span, mn_p1, one_m_mx;
    static __device__ __forceinline__ PixelMap make(int mode, int clamp, float mn, float mx, const MapConst& k) {
        PixelMap m;
        m.mode = mode, m.clamp = clamp, m.mn = mn, m.mx = mx;
        m.a = k.a, m.two_a = k.two_a, m.a_m_1 = k.a_m_1, m.one_m_a = k.one_m_a;
        m.span = __fsub_rn(mx, mn), m.mn_p1 = __fadd_rn(mn, 1.f), m.one_m_mx = __fsub_rn(1.f, mx);
        return m;
    }
    template <int MODE>
    __device__ __forceinline__ float at(float x) const {
        if (clamp) x = fminf(fmaxf(x, -1.f), 1.f);
        if (MODE == MOBI_RANGE_MAP_DEPTH_UNNORM) {
            const bool mid = x >= -a && x <= a, low = x >= -1.f && x < -a, high = x > a && x <= 1.f;
            if (!(mid || low || high)) return x;
            float t = __fadd_rn(x, mid ? a : (low ? 1.f : -a));  // x - a == x + (-a) exactly
            t = low ? -t : t;
            const float u = __fmul_rn(t, mid ? span : (low ? mn_p1 : one_m_mx));
            return __fadd_rn(mid ? mn : (low ? -1.f : mx), __fdiv_rn(u, mid ? two_a : (low ? a_m_1 : one_m_a)));
        }
        if (MODE == MOBI_RANGE_MAP_DEPTH_NORM) return depth_norm_fwd(x, mn, mx, a, two_a, one_m_a);
        if (MODE == MOBI_RANGE_MAP_INT_UNNORM) return int_unnorm(x);
        return x;
    }
};

// Sum of a kh x KW window in the order F.avg_pool2d's CPU kernel uses (row by row, one fp32 accumulator).  The adds keep
// that order; the loads do not depend on them: a row is one or two 16-byte loads and four rows are in flight per thread.
template <int KW, int MODE>
__device__ __forceinline__ float window_sum(const float* __restrict__ win, int kh, int cw, int kw_rt, bool vec, const PixelMap& m) {
    float acc = 0.f;
    if (KW >= 4 && vec) {
#pragma unroll 4
        for (int i = 0; i < kh; ++i) {
            const float4* row = reinterpret_cast<const float4*>(win + (long long)i * cw);
#pragma unroll
            for (int q = 0; q < KW / 4; ++q) {
                const float4 v = __ldg(row + q);
                acc = __fadd_rn(acc, m.at<MODE>(v.x));
                acc = __fadd_rn(acc, m.at<MODE>(v.y));
                acc = __fadd_rn(acc, m.at<MODE>(v.z));
                acc = __fadd_rn(acc, m.at<MODE>(v.w));
            }
        }
    } else if (KW == 2 && vec) {
#pragma unroll 4
        for (int i = 0; i < kh; ++i) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(win + (long long)i * cw));
            acc = __fadd_rn(acc, m.at<MODE>(v.x));
            acc = __fadd_rn(acc, m.at<MODE>(v.y));
        }
    } else {
        const int kw = KW > 0 ? KW : kw_rt;
#pragma unroll 4
        for (int i = 0; i < kh; ++i) {
            const float* row = win + (long long)i * cw;
            for (int q = 0; q < kw; ++q) acc = __fadd_rn(acc, m.at<MODE>(__ldg(row + q)));
        }
    }
    return acc;
}

template <int MODE>
__device__ __forceinline__ float window_sum_kw(const float* __restrict__ win, int kh, int kw, int cw, bool vec, const PixelMap& m) {
    switch (kw) {
        case 1: return window_sum<1, MODE>(win, kh, cw, kw, vec, m);
        case 2: return window_sum<2, MODE>(win, kh, cw, kw, vec, m);
        case 4: return window_sum<4, MODE>(win, kh, cw, kw, vec, m);
        case 8: return window_sum<8, MODE>(win, kh, cw, kw, vec, m);
        default: return window_sum<0, MODE>(win, kh, cw, kw, false, m);
    }
}

// ---- undo_default_transforms ------------------------------------------------------------------------------------------
struct UndoParams {
    const float* crop[2];
    const float* orig[2];
    float* out[2];
    const float* min_d;
    const float* max_d;
    const long long* crop_left;
    const long long* width_crop;
    long long crop_bs, orig_bs, out_bs;
    int channels, ch, cw, H, W, zero_context, clamp, vec4;
    int map[2];
    MapConst k;
};

__global__ void __launch_bounds__(256) range_undo_kernel(UndoParams p) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int y = blockIdx.x / ((p.W + 255) / 256);
    const int x = (blockIdx.x % ((p.W + 255) / 256)) * 256 + threadIdx.x;
    if (x >= p.W) return;
    const long long o = (long long)y * p.W + x;
    float base = p.orig[c][b * p.orig_bs + o];
    if (p.zero_context && c == 0) base = __fsub_rn(__fmul_rn(base, 0.f), 1.f);  // range_depth_orig * 0 - 1 (utils.py:486-487)
    const int wc = (int)p.width_crop[b];
    long long cl = p.crop_left[b] % p.W;
    if (cl < 0) cl += p.W;  // Python's % is non-negative
    int j = x - (int)cl;
    if (j < 0) j += p.W;
    float v = base;
    if (j < wc && wc > 0) {
        const float* src = p.crop[c] + b * p.crop_bs;
        const float mn = p.min_d ? p.min_d[b] : 0.f, mx = p.max_d ? p.max_d[b] : 0.f;
        const int mode = p.map[c];
        if (p.ch == p.H && p.cw == wc) {  // already at the target size (lidar_converter.py:259)
            v = apply_map(src[(long long)y * p.cw + j], mode, p.clamp, mn, mx, p.k);
        } else if (p.ch % p.H == 0 && p.cw % wc == 0) {
            // F.avg_pool2d: the window is summed row by row into ONE fp32 accumulator, then divided once
            const int kh = p.ch / p.H, kw = p.cw / wc;
            const float* win = src + (long long)y * kh * p.cw + (long long)j * kw;
            const PixelMap m = PixelMap::make(mode, p.clamp, mn, mx, p.k);
            float acc;
            switch (mode) {
                case MOBI_RANGE_MAP_DEPTH_UNNORM: acc = window_sum_kw<MOBI_RANGE_MAP_DEPTH_UNNORM>(win, kh, kw, p.cw, p.vec4, m); break;
                case MOBI_RANGE_MAP_DEPTH_NORM: acc = window_sum_kw<MOBI_RANGE_MAP_DEPTH_NORM>(win, kh, kw, p.cw, p.vec4, m); break;
                case MOBI_RANGE_MAP_INT_UNNORM: acc = window_sum_kw<MOBI_RANGE_MAP_INT_UNNORM>(win, kh, kw, p.cw, p.vec4, m); break;
                default: acc = window_sum_kw<MOBI_RANGE_MAP_NONE>(win, kh, kw, p.cw, p.vec4, m); break;
            }
            v = __fdiv_rn(acc, (float)(kh * kw));
        } else {
            // cv2.resize INTER_NEAREST: source index = min(floor(dst * (1 / (dst_size / src_size))), src_size - 1), doubles
            const double ifx = 1.0 / ((double)wc / (double)p.cw), ify = 1.0 / ((double)p.H / (double)p.ch);
            const int sx = min((int)floor(j * ifx), p.cw - 1), sy = min((int)floor(y * ify), p.ch - 1);
            v = apply_map(src[(long long)sy * p.cw + sx], mode, p.clamp, mn, mx, p.k);
        }
        if (v == -1000.f) v = base;  // the `ignore` sentinel of lidar_converter.py:452, 466
    }
    p.out[c][b * p.out_bs + o] = v;
}

// ---- range2pcd / save_samples composite -------------------------------------------------------------------------------
constexpr int RC_CLUSTER = 8;
constexpr int RC_THREADS = 512;
constexpr int RC_WARPS = RC_THREADS / 32;
constexpr int RC_MAX_WORDS = 2048;  // 32-pixel ballot words per CTA: sweeps of up to 8 * 2048 * 32 = 524,288 pixels

struct CloudParams {
    const float* depth;  // plain mode: range depth; composite: the generated (un-cropped) sweep
    const float* label;  // plain: optional value carried with each point; composite: generated intensity
    const float* depth_orig;
    const float* int_orig;
    const float* gt_mask;
    const float* bbox;  // [batch, 8, 3]
    const float* pitch;
    const float* yaw;
    float* range_pred;     // composite out [batch, 4, P]
    uint8_t* pred_mask;    // composite out [batch, P]
    float* points;         // out [batch, P, point_stride]
    int32_t* index;        // optional out [batch, P]
    int32_t* count;        // out [batch]
    int P, W, H, point_stride;
    float depth_min, depth_max, depth_scale;
};

__device__ __forceinline__ float range_metres(float d, float scale) {
    return __fmul_rn(__fmul_rn(__fadd_rn(d, 1.f), 0.5f), scale);  // ((d + 1) / 2) * depth_interval[1]
}

template <bool COMPOSITE>
__global__ void __cluster_dims__(RC_CLUSTER, 1, 1) __launch_bounds__(RC_THREADS) range_cloud_kernel(CloudParams p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t words[RC_MAX_WORDS];
    __shared__ int word_off[RC_MAX_WORDS];
    __shared__ float plane[6][4];
    __shared__ int cta_total;
    __shared__ int warp_sums[RC_WARPS];

    const int b = blockIdx.x / RC_CLUSTER;
    const unsigned rank = cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // every CTA owns a contiguous, 32-aligned run of pixels so that (rank, word, lane) order is the row-major pixel order
    const int words_total = (p.P + 31) / 32;
    const int words_per_cta = (words_total + RC_CLUSTER - 1) / RC_CLUSTER;
    const int w_beg = min((int)rank * words_per_cta, words_total), w_end = min(w_beg + words_per_cta, words_total);
    const int nwords = w_end - w_beg;
    const long long sample = (long long)b * p.P;

    if (COMPOSITE) {
        // corner_to_surfaces_3d + surface_equ_3d: inward normals n = (p0 - p1) x (p1 - p2), d = -n . p0, fp32
        if (threadIdx.x < 6) {
            const int table[6][3] = {{0, 1, 2}, {7, 6, 5}, {0, 3, 7}, {1, 5, 6}, {0, 4, 5}, {3, 2, 6}};
            const float* c = p.bbox + (long long)b * 24;
            const float* p0 = c + 3 * table[threadIdx.x][0];
            const float* p1 = c + 3 * table[threadIdx.x][1];
            const float* p2 = c + 3 * table[threadIdx.x][2];
            const float ax = __fsub_rn(p0[0], p1[0]), ay = __fsub_rn(p0[1], p1[1]), az = __fsub_rn(p0[2], p1[2]);
            const float bx = __fsub_rn(p1[0], p2[0]), by = __fsub_rn(p1[1], p2[1]), bz = __fsub_rn(p1[2], p2[2]);
            const float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
            const float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
            const float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(nx, p0[0]), __fmul_rn(ny, p0[1])), __fmul_rn(nz, p0[2]));
            plane[threadIdx.x][0] = nx;
            plane[threadIdx.x][1] = ny;
            plane[threadIdx.x][2] = nz;
            plane[threadIdx.x][3] = -d;
        }
        __syncthreads();
    }

    // pass 1: per pixel decisions; validity ballots in pixel order
    for (int w = warp; w < nwords; w += RC_WARPS) {
        const int px = (w_beg + w) * 32 + lane;
        bool valid = false;
        if (px < p.P) {
            float d = p.depth[sample + px];
            if (COMPOSITE) {
                const float pitch = p.pitch[sample + px], yaw = p.yaw[sample + px];
                const float m = range_metres(d, p.depth_scale);
                bool inside = false;
                if (m > p.depth_min && m < p.depth_max) {  // only kept points are tested (inference_test_bench.py:588-594)
                    float sy, cy, sp, cp;
                    sincosf(yaw, &sy, &cy);
                    sincosf(pitch, &sp, &cp);
                    const float x = __fmul_rn(__fmul_rn(cy, cp), m), y = __fmul_rn(__fmul_rn(-sy, cp), m), z = __fmul_rn(sp, m);
                    inside = true;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, plane[k][0]), __fmul_rn(y, plane[k][1])),
                                                            __fmul_rn(z, plane[k][2])), plane[k][3]);
                        if (s >= 0.f) inside = false;
                    }
                }
                const bool paste = inside || (p.gt_mask[sample + px] != 0.f);
                const float inten = paste ? p.label[sample + px] : p.int_orig[sample + px];
                d = paste ? d : p.depth_orig[sample + px];
                float* rp = p.range_pred + (long long)b * 4 * p.P + px;
                rp[0] = d;
                rp[p.P] = inten;
                rp[2LL * p.P] = pitch;
                rp[3LL * p.P] = yaw;
                p.pred_mask[sample + px] = inside ? 1 : 0;
            }
            const float m = range_metres(d, p.depth_scale);
            valid = m > p.depth_min && m < p.depth_max;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) words[w] = bal;
    }
    __syncthreads();

    // exclusive scan of the word populations (block-wide, nwords <= RC_MAX_WORDS)
    {
        int carry = 0;
        for (int base = 0; base < nwords; base += RC_THREADS) {
            const int w = base + threadIdx.x;
            const int c = w < nwords ? __popc(words[w]) : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            int before = carry;
            for (int q = 0; q < warp; ++q) before += warp_sums[q];
            if (w < nwords) word_off[w] = before + incl - c;
            int tile_total = 0;
            for (int q = 0; q < RC_WARPS; ++q) tile_total += warp_sums[q];
            carry += tile_total;
            __syncthreads();
        }
        if (threadIdx.x == 0) cta_total = carry;
    }
    cluster.sync();
    int cta_base = 0, sweep_total = 0;
    for (unsigned r = 0; r < RC_CLUSTER; ++r) {
        const int t = *cluster.map_shared_rank(&cta_total, r);
        if (r < rank) cta_base += t;
        sweep_total += t;
    }
    cluster.sync();  // cta_total stays alive until every CTA of the cluster has read it
    if (rank == 0 && threadIdx.x == 0) p.count[b] = sweep_total;

    // pass 2: write the kept points at their ordered positions
    for (int w = warp; w < nwords; w += RC_WARPS) {
        const uint32_t bal = words[w];
        if (!((bal >> lane) & 1u)) continue;
        const int px = (w_beg + w) * 32 + lane;
        const int pos = cta_base + word_off[w] + __popc(bal & ((1u << lane) - 1u));
        float d, lab = 0.f;
        if (COMPOSITE) {
            const float* rp = p.range_pred + (long long)b * 4 * p.P + px;
            d = rp[0];
            lab = rp[p.P];
        } else {
            d = p.depth[sample + px];
            if (p.label) lab = p.label[sample + px];
        }
        const float m = range_metres(d, p.depth_scale);
        float sy, cy, sp, cp;
        sincosf(p.yaw[sample + px], &sy, &cy);
        sincosf(p.pitch[sample + px], &sp, &cp);
        float* out = p.points + (sample + pos) * p.point_stride;
        out[0] = __fmul_rn(__fmul_rn(cy, cp), m);
        out[1] = __fmul_rn(__fmul_rn(-sy, cp), m);
        out[2] = __fmul_rn(sp, m);
        if (p.point_stride > 3) out[3] = lab;
        if (p.point_stride > 4) out[4] = (float)(p.H - 1 - px / p.W);  // beam index (lidar_converter.py:170-174)
        if (p.index) p.index[sample + pos] = px;
    }
}

// ---- points_in_bbox_corners --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
points_in_boxes_kernel(const float* __restrict__ pts, int pt_stride, const float* __restrict__ corners, uint8_t* __restrict__ out,
                       int n, int m) {
    extern __shared__ float planes[];  // [m][6][4]
    for (int i = threadIdx.x; i < m * 6; i += blockDim.x) {
        const int table[6][3] = {{0, 1, 2}, {7, 6, 5}, {0, 3, 7}, {1, 5, 6}, {0, 4, 5}, {3, 2, 6}};
        const int j = i / 6, k = i % 6;
        const float* c = corners + (long long)j * 24;
        const float* p0 = c + 3 * table[k][0];
        const float* p1 = c + 3 * table[k][1];
        const float* p2 = c + 3 * table[k][2];
        const float ax = __fsub_rn(p0[0], p1[0]), ay = __fsub_rn(p0[1], p1[1]), az = __fsub_rn(p0[2], p1[2]);
        const float bx = __fsub_rn(p1[0], p2[0]), by = __fsub_rn(p1[1], p2[1]), bz = __fsub_rn(p1[2], p2[2]);
        const float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
        const float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
        const float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
        planes[i * 4 + 0] = nx;
        planes[i * 4 + 1] = ny;
        planes[i * 4 + 2] = nz;
        planes[i * 4 + 3] = -__fadd_rn(__fadd_rn(__fmul_rn(nx, p0[0]), __fmul_rn(ny, p0[1])), __fmul_rn(nz, p0[2]));
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = pts[(long long)i * pt_stride], y = pts[(long long)i * pt_stride + 1], z = pts[(long long)i * pt_stride + 2];
        for (int j = 0; j < m; ++j) {
            bool inside = true;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float* pl = planes + (j * 6 + k) * 4;
                const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, pl[0]), __fmul_rn(y, pl[1])), __fmul_rn(z, pl[2])), pl[3]);
                if (s >= 0.f) inside = false;
            }
            out[(long long)i * m + j] = inside ? 1 : 0;
        }
    }
}

}  // namespace mobi

using namespace mobi;

#define MOBI_STREAM cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_)

extern "C" int mobi_range_map(const mobi_range_map_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->in && a->out && a->batch > 0 && a->n > 0, "mobi_range_map: bad argument");
    MOBI_CHECK(a->mode >= MOBI_RANGE_MAP_NONE && a->mode <= MOBI_RANGE_MAP_INT_UNNORM, "mobi_range_map: unknown mode %d", a->mode);
    const bool depth = a->mode == MOBI_RANGE_MAP_DEPTH_NORM || a->mode == MOBI_RANGE_MAP_DEPTH_UNNORM;
    MOBI_CHECK(!depth || (a->min_d && a->max_d), "mobi_range_map: the depth maps need min_d and max_d");
    MOBI_CHECK(!depth || (a->alpha > 0.0 && a->alpha <= 1.0), "mobi_range_map: alpha must be in the range 0 to 1");
    long long bx = (a->n + 255) / 256;
    const long long cap = std::max(1, 4 * sm_count() / a->batch);
    if (bx > cap) bx = cap;
    range_map_kernel<<<dim3((unsigned)bx, (unsigned)a->batch), 256, 0, stream>>>(
        a->in, a->out, a->min_d, a->max_d, a->n, a->in_stride ? a->in_stride : a->n, a->out_stride ? a->out_stride : a->n,
        a->mode, a->clamp_input, map_const(depth ? a->alpha : 0.75));
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_range_undo_transforms(const mobi_range_undo_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->batch > 0 && (a->channels == 1 || a->channels == 2), "mobi_range_undo_transforms: bad argument");
    MOBI_CHECK(a->crop_left && a->width_crop && a->crop_h > 0 && a->crop_w > 0 && a->H > 0 && a->W > 0,
               "mobi_range_undo_transforms: bad argument");
    UndoParams p{};
    bool need_minmax = false;
    for (int c = 0; c < a->channels; ++c) {
        MOBI_CHECK(a->crop[c] && a->orig[c] && a->out[c], "mobi_range_undo_transforms: channel %d has a null pointer", c);
        MOBI_CHECK(a->map[c] >= MOBI_RANGE_MAP_NONE && a->map[c] <= MOBI_RANGE_MAP_INT_UNNORM,
                   "mobi_range_undo_transforms: unknown map %d", a->map[c]);
        p.crop[c] = a->crop[c];
        p.orig[c] = a->orig[c];
        p.out[c] = a->out[c];
        p.map[c] = a->map[c];
        need_minmax |= a->map[c] == MOBI_RANGE_MAP_DEPTH_NORM || a->map[c] == MOBI_RANGE_MAP_DEPTH_UNNORM;
    }
    MOBI_CHECK(!need_minmax || (a->min_d && a->max_d && a->alpha > 0.0 && a->alpha <= 1.0),
               "mobi_range_undo_transforms: the fused depth map needs min_d, max_d and alpha in (0, 1]");
    p.min_d = a->min_d;
    p.max_d = a->max_d;
    p.crop_left = reinterpret_cast<const long long*>(a->crop_left);
    p.width_crop = reinterpret_cast<const long long*>(a->width_crop);
    p.crop_bs = a->crop_batch_stride ? a->crop_batch_stride : (long long)a->crop_h * a->crop_w;
    p.orig_bs = a->orig_batch_stride ? a->orig_batch_stride : (long long)a->H * a->W;
    p.out_bs = a->out_batch_stride ? a->out_batch_stride : (long long)a->H * a->W;
    p.channels = a->channels;
    p.ch = a->crop_h;
    p.cw = a->crop_w;
    p.H = a->H;
    p.W = a->W;
    p.zero_context = a->zero_context;
    p.clamp = a->clamp_input;
    p.k = map_const(need_minmax ? a->alpha : 0.75);
    p.vec4 = (p.cw % 4 == 0) && (p.crop_bs % 4 == 0);
    for (int c = 0; c < a->channels; ++c) p.vec4 = p.vec4 && (reinterpret_cast<uintptr_t>(a->crop[c]) % 16 == 0);
    const unsigned gx = (unsigned)(a->H * ((a->W + 255) / 256));
    range_undo_kernel<<<dim3(gx, (unsigned)a->channels, (unsigned)a->batch), 256, 0, stream>>>(p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

static int cloud_common(CloudParams& p, int batch, int H, int W, int point_stride, float depth_min, float depth_max,
                        const char* who) {
    MOBI_CHECK(batch > 0 && H > 0 && W > 0, "%s: bad shape", who);
    MOBI_CHECK((long long)H * W <= (long long)RC_CLUSTER * RC_MAX_WORDS * 32, "%s: sweeps of more than %d pixels are not supported",
               who, RC_CLUSTER * RC_MAX_WORDS * 32);
    MOBI_CHECK(point_stride >= 3, "%s: point_stride must be >= 3", who);
    MOBI_CHECK(depth_max > depth_min, "%s: empty depth interval", who);
    p.P = H * W;
    p.W = W;
    p.H = H;
    p.point_stride = point_stride;
    p.depth_min = depth_min;
    p.depth_max = depth_max;
    p.depth_scale = depth_max;
    return 0;
}

extern "C" int mobi_range2pcd(const mobi_range2pcd_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->depth && a->pitch && a->yaw && a->points && a->count, "mobi_range2pcd: bad argument");
    CloudParams p{};
    if (cloud_common(p, a->batch, a->H, a->W, a->point_stride, a->depth_min, a->depth_max, "mobi_range2pcd")) return 1;
    p.depth = a->depth;
    p.label = a->label;
    p.pitch = a->pitch;
    p.yaw = a->yaw;
    p.points = a->points;
    p.index = a->index;
    p.count = a->count;
    range_cloud_kernel<false><<<(unsigned)(a->batch * RC_CLUSTER), RC_THREADS, 0, stream>>>(p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_range_composite(const mobi_range_composite_args* a, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(a && a->sample_depth && a->sample_int && a->depth_orig && a->int_orig && a->gt_mask && a->bbox && a->pitch &&
                   a->yaw && a->range_pred && a->pred_mask && a->points && a->count,
               "mobi_range_composite: bad argument");
    CloudParams p{};
    if (cloud_common(p, a->batch, a->H, a->W, 5, a->depth_min, a->depth_max, "mobi_range_composite")) return 1;
    p.depth = a->sample_depth;
    p.label = a->sample_int;
    p.depth_orig = a->depth_orig;
    p.int_orig = a->int_orig;
    p.gt_mask = a->gt_mask;
    p.bbox = a->bbox;
    p.pitch = a->pitch;
    p.yaw = a->yaw;
    p.range_pred = a->range_pred;
    p.pred_mask = a->pred_mask;
    p.points = a->points;
    p.index = nullptr;
    p.count = a->count;
    range_cloud_kernel<true><<<(unsigned)(a->batch * RC_CLUSTER), RC_THREADS, 0, stream>>>(p);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_points_in_boxes(const float* points, int32_t point_stride, const float* corners, uint8_t* out, int32_t n,
                                    int32_t m, void* stream_) {
    MOBI_STREAM;
    MOBI_CHECK(points && corners && out && n >= 0 && m > 0 && point_stride >= 3, "mobi_points_in_boxes: bad argument");
    MOBI_CHECK(m <= 1024, "mobi_points_in_boxes: at most 1024 boxes per call");
    if (n == 0) return 0;
    const int blocks = std::min((n + 255) / 256, 8 * sm_count());
    points_in_boxes_kernel<<<blocks, 256, (size_t)m * 96, stream>>>(points, point_stride, corners, out, n, m);
    MOBI_CUDA(cudaGetLastError());
    return 0;
}
