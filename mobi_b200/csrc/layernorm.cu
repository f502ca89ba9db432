// LayerNorm family of the transformer blocks (HBM-bound; one warp owns one token row held in registers).
//
//   layernorm_kernel      LN -> bf16 with segment gather and fused broadcast add (kept from round 1, 32-bit index math)
//   ln_dual_kernel        ONE pass over all rows of the interleaved batch: even (camera) and odd (lidar) batch rows each
//                         get their own action (LayerNorm with their own affine, or a plain bf16 cast) and their own
//                         compact output, replacing two gathers (attention.py:246-261)
//   ln_adapter_kernel     the bbox/reference adapter (attention.py:237-243) fused with the LayerNorms around it:
//                           x += vec2                                   (attn2, one key => constant vector per row)
//                           d  = LN_stats(x)                            (cond_adapter_norm)
//                           s_j = rstd * <d * gamma, U_j> + <beta, U_j> (U = W_q^T k * scale, 2 keys x 8 heads = 16 rows)
//                           p   = 2-way softmax per head ;  x += sum_j p_j Z_j + zb     (Z = W_conn W_out v)
//                         followed by the same dual action as ln_dual_kernel on the updated row.
//                         The softmax is over TWO keys per head, so only score differences matter:
//                           p_h = sigmoid(<d, U_h0 - U_h1> * rstd + (sb_h0 - sb_h1)),
//                           x  += (zb + sum_h Z_h1) + sum_h p_h (Z_h0 - Z_h1):
//                         8 difference rows per table instead of 16 rows (built while the tables are staged in shared
//                         memory as fp32), which halves the shared-memory reads, the FMAs and the accumulators and lets
//                         four token rows per warp share every table load.
#include "../../include/mobi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace mobi {

constexpr int LN_MAXV = 12;  // float4 vectors per lane: C <= 12*128 = 1536 (kernels are instantiated for 3, 6, 12)

template <int MAXV>
__device__ __forceinline__ void ln_load_row(const float* __restrict__ xr, int nv, int lane, float4 (&v)[MAXV]) {
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) v[k] = reinterpret_cast<const float4*>(xr)[i];
        else v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int MAXV>
__device__ __forceinline__ void ln_stats(const float4 (&v)[MAXV], int nv, int lane, int C, float eps, float& mean,
                                         float& rstd) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);  // padding vectors are zero
    mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        if (lane + 32 * k < nv) {
            const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
    }
    rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
}

// mode 1: LayerNorm with (gamma, beta) -> bf16 ; mode 2: plain cast -> bf16
template <int MAXV>
__device__ __forceinline__ void ln_emit(const float4 (&v)[MAXV], int nv, int lane, int C, float eps, int mode,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        __nv_bfloat16* __restrict__ orow_) {
    uint2* orow = reinterpret_cast<uint2*>(orow_);
    if (mode == 2) {
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) orow[i] = make_uint2(pack_bf16x2(v[k].x, v[k].y), pack_bf16x2(v[k].z, v[k].w));
        }
        return;
    }
    float mean, rstd;
    ln_stats(v, nv, lane, C, eps, mean, rstd);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i);
            const float y0 = (v[k].x - mean) * rstd * g.x + b.x;
            const float y1 = (v[k].y - mean) * rstd * g.y + b.y;
            const float y2 = (v[k].z - mean) * rstd * g.z + b.z;
            const float y3 = (v[k].w - mean) * rstd * g.w + b.w;
            orow[i] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
        }
    }
}

template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ out, const float* __restrict__ add_vec, int rows, int C, int seg,
                 int seg_stride, int seg_offset, int add_rows_per_vec, float eps) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int sg = row / seg;
    const long long in_row = (long long)sg * seg_stride + seg_offset + (row - sg * seg);
    float* xr = x + in_row * C;
    const int nv = C >> 2;
    float4 v[MAXV];
    ln_load_row(xr, nv, lane, v);
    if (add_vec) {
        const float4* av = reinterpret_cast<const float4*>(add_vec + (in_row / add_rows_per_vec) * C);
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) {
                const float4 a = __ldg(av + i);
                v[k].x += a.x;
                v[k].y += a.y;
                v[k].z += a.z;
                v[k].w += a.w;
                reinterpret_cast<float4*>(xr)[i] = v[k];
            }
        }
    }
    ln_emit(v, nv, lane, C, eps, gamma ? 1 : 2, gamma, beta, out + (long long)row * C);
}

struct DualParams {
    const float* gamma[2];
    const float* beta[2];
    __nv_bfloat16* out[2];
    int mode[2];  // 0 = nothing, 1 = LayerNorm, 2 = cast
    int pair;     // 1: batch rows alternate (even -> slot 0, odd -> slot 1) and outputs are compacted per slot
};

// Output row of token `t` of batch row `b` for its slot.
template <int MAXV>
__device__ __forceinline__ void dual_emit(const DualParams& d, const float4 (&v)[MAXV], int nv, int lane, int C,
                                          float eps, int b, int t, int T) {
    const int slot = d.pair ? (b & 1) : 0;
    const int mode = d.mode[slot];
    if (mode == 0) return;
    const long long orow = d.pair ? ((long long)(b >> 1) * T + t) : ((long long)b * T + t);
    ln_emit(v, nv, lane, C, eps, mode, d.gamma[slot], d.beta[slot], d.out[slot] + orow * C);
}

template <int MAXV>
__global__ void __launch_bounds__(256)
ln_dual_kernel(const float* __restrict__ x, DualParams d, int rows, int C, int T, float eps) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = row / T, t = row - b * T;
    if (d.mode[d.pair ? (b & 1) : 0] == 0) return;
    const int nv = C >> 2;
    float4 v[MAXV];
    ln_load_row(x + (long long)row * C, nv, lane, v);
    dual_emit(d, v, nv, lane, C, eps, b, t, T);
}

// ------------------------------------------------------------------------------------------------ adapter
constexpr int AD_HK = 16;      // table rows of the ABI: 2 keys x 8 head slots (heads < 8 are zero padded by the host)
constexpr int AD_H = 8;        // difference rows kept in shared memory (key 0 minus key 1 of every head slot)
constexpr int AD_WARPS = 8;
#ifndef AD_B3
#define AD_B3 3   // CTAs per SM the register allocation must allow, by channel-count class (measured: tools/gpu_r02_lnad.sh)
#endif
#ifndef AD_B6
#define AD_B6 2
#endif
#ifndef AD_B12
#define AD_B12 2
#endif
#ifndef AD_R3
#define AD_R3 2   // token rows per warp per iteration (share the table loads), by channel-count class
#endif
#ifndef AD_R6
#define AD_R6 2
#endif
#ifndef AD_R12
#define AD_R12 1
#endif

struct AdapterParams {
    float* x;
    const float* add_vec;   // [R, C] or null
    const float* gamma;     // cond_adapter_norm
    const float* beta;
    const float* Ug;        // [R, 16, C]  = U * gamma
    const float* sb;        // [R, 16]     = <beta, U_j>
    const float* Z;         // [R, 16, C]
    const float* zb;        // [C]
    int T, C, rows_per_cta;
    float eps;
};

// Sums 8 per-lane values over the 32 lanes with 9 shuffles; on return lane l holds the total of slot
// h(l) = 4*bit4(l) + 2*bit3(l) + bit2(l)   (the four lanes that share bits 4..2 hold the same slot).
__device__ __forceinline__ float reduce8(float (&a)[AD_H], int lane) {
    float b4[4], b2[2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float keep = hi ? a[i + 4] : a[i];
            const float send = hi ? a[i] : a[i + 4];
            b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float keep = hi ? b4[i + 2] : b4[i];
            const float send = hi ? b4[i] : b4[i + 2];
            b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    const bool hi = lane & 4;
    const float keep = hi ? b2[1] : b2[0];
    const float send = hi ? b2[0] : b2[1];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// grid (ceil(T / rows_per_cta), R); 256 threads.  dynamic smem: dU[8][C] | dZ[8][C] | zc[C] fp32
template <int MAXV, int AD_ROWS, int MINB>
__global__ void __launch_bounds__(AD_WARPS * 32, MINB)
ln_adapter_kernel(AdapterParams p, DualParams d) {
    extern __shared__ float ad_smem[];
    pdl_wait();   // (the tables read below were written once per sampling run, but x comes from the previous kernel)
    pdl_trigger();
    const int C = p.C, T = p.T;
    const int nv = C >> 2;
    const int b = blockIdx.y;
    float4* sU = reinterpret_cast<float4*>(ad_smem);
    float4* sZ = sU + AD_H * nv;
    float4* sZc = sZ + AD_H * nv;
    {
        const float4* gU = reinterpret_cast<const float4*>(p.Ug + (long long)b * AD_HK * C);
        const float4* gZ = reinterpret_cast<const float4*>(p.Z + (long long)b * AD_HK * C);
        const float4* zb4 = reinterpret_cast<const float4*>(p.zb);
        for (int i = threadIdx.x; i < AD_H * nv; i += blockDim.x) {
            const float4 u0 = __ldg(gU + i), u1 = __ldg(gU + AD_H * nv + i);       // key 0 / key 1 of the same head slot
            const float4 z0 = __ldg(gZ + i), z1 = __ldg(gZ + AD_H * nv + i);
            sU[i] = make_float4(u0.x - u1.x, u0.y - u1.y, u0.z - u1.z, u0.w - u1.w);
            sZ[i] = make_float4(z0.x - z1.x, z0.y - z1.y, z0.z - z1.z, z0.w - z1.w);
        }
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            float4 acc = __ldg(zb4 + i);
#pragma unroll
            for (int h = 0; h < AD_H; ++h) {
                const float4 z1 = __ldg(gZ + (AD_H + h) * nv + i);
                acc.x += z1.x; acc.y += z1.y; acc.z += z1.z; acc.w += z1.w;
            }
            sZc[i] = acc;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = 4 * ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1) + ((lane >> 2) & 1);
    const float my_dsb = __ldg(p.sb + b * AD_HK + slot) - __ldg(p.sb + b * AD_HK + AD_H + slot);
    const int t_begin = blockIdx.x * p.rows_per_cta;
    const int t_end = min(T, t_begin + p.rows_per_cta);
    const float4* av = p.add_vec ? reinterpret_cast<const float4*>(p.add_vec + (long long)b * C) : nullptr;

    for (int t0 = t_begin + warp * AD_ROWS; t0 < t_end; t0 += AD_WARPS * AD_ROWS) {
        float4 v[AD_ROWS][MAXV];
        bool live[AD_ROWS];
#pragma unroll
        for (int r = 0; r < AD_ROWS; ++r) {
            live[r] = (t0 + r) < t_end;
            const int t = live[r] ? t0 + r : t0;
            ln_load_row(p.x + ((long long)b * T + t) * C, nv, lane, v[r]);
        }
        if (av) {
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int i = lane + 32 * k;
                if (i < nv) {
                    const float4 a = __ldg(av + i);
#pragma unroll
                    for (int r = 0; r < AD_ROWS; ++r) {
                        v[r][k].x += a.x;
                        v[r][k].y += a.y;
                        v[r][k].z += a.z;
                        v[r][k].w += a.w;
                    }
                }
            }
        }
        float mean[AD_ROWS], rstd[AD_ROWS];
#pragma unroll
        for (int r = 0; r < AD_ROWS; ++r) ln_stats(v[r], nv, lane, C, p.eps, mean[r], rstd[r]);
        // score differences against the 8 table rows (centred values, so no cancellation against the mean)
        float acc[AD_ROWS][AD_H];
#pragma unroll
        for (int r = 0; r < AD_ROWS; ++r)
#pragma unroll
            for (int j = 0; j < AD_H; ++j) acc[r][j] = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) {
                float4 dv[AD_ROWS];
#pragma unroll
                for (int r = 0; r < AD_ROWS; ++r)
                    dv[r] = make_float4(v[r][k].x - mean[r], v[r][k].y - mean[r], v[r][k].z - mean[r], v[r][k].w - mean[r]);
#pragma unroll
                for (int j = 0; j < AD_H; ++j) {
                    const float4 u = sU[j * nv + i];
#pragma unroll
                    for (int r = 0; r < AD_ROWS; ++r)
                        acc[r][j] += dv[r].x * u.x + dv[r].y * u.y + dv[r].z * u.z + dv[r].w * u.w;
                }
            }
        }
        float pj[AD_ROWS][AD_H];
#pragma unroll
        for (int r = 0; r < AD_ROWS; ++r) {
            const float ds = reduce8(acc[r], lane) * rstd[r] + my_dsb;   // this lane's head slot: s(key 0) - s(key 1)
            const float pr = 1.0f / (1.0f + __expf(-ds));                // 2-way softmax: weight of key 0
#pragma unroll
            for (int j = 0; j < AD_H; ++j) {
                // a lane that holds slot j: bits 4..2 of the lane id spell j
                const int src = ((j >> 2) & 1) * 16 + ((j >> 1) & 1) * 8 + (j & 1) * 4;
                pj[r][j] = __shfl_sync(0xffffffffu, pr, src);
            }
        }
#pragma unroll
        for (int k = 0; k < MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) {
                const float4 zc = sZc[i];
                float4 o[AD_ROWS];
#pragma unroll
                for (int r = 0; r < AD_ROWS; ++r) o[r] = zc;
#pragma unroll
                for (int j = 0; j < AD_H; ++j) {
                    const float4 z = sZ[j * nv + i];
#pragma unroll
                    for (int r = 0; r < AD_ROWS; ++r) {
                        o[r].x += pj[r][j] * z.x;
                        o[r].y += pj[r][j] * z.y;
                        o[r].z += pj[r][j] * z.z;
                        o[r].w += pj[r][j] * z.w;
                    }
                }
#pragma unroll
                for (int r = 0; r < AD_ROWS; ++r) {
                    v[r][k].x += o[r].x;
                    v[r][k].y += o[r].y;
                    v[r][k].z += o[r].z;
                    v[r][k].w += o[r].w;
                    if (live[r])
                        reinterpret_cast<float4*>(p.x + ((long long)b * T + t0 + r) * C)[i] = v[r][k];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < AD_ROWS; ++r)
            if (live[r]) dual_emit(d, v[r], nv, lane, C, p.eps, b, t0 + r, T);
    }
}

static int fill_dual(DualParams& d, const mobi_ln_dual_spec* s) {
    d.pair = s->pair;
    for (int i = 0; i < 2; ++i) {
        d.gamma[i] = s->gamma[i];
        d.beta[i] = s->beta[i];
        d.out[i] = reinterpret_cast<__nv_bfloat16*>(s->out[i]);
        d.mode[i] = s->mode[i];
        MOBI_CHECK(d.mode[i] >= 0 && d.mode[i] <= 2, "ln dual: bad mode %d", d.mode[i]);
        MOBI_CHECK(d.mode[i] == 0 || d.out[i] != nullptr, "ln dual: slot %d has no output", i);
        MOBI_CHECK(d.mode[i] != 1 || (d.gamma[i] && d.beta[i]), "ln dual: slot %d LayerNorm needs gamma/beta", i);
    }
    return 0;
}

}  // namespace mobi

using namespace mobi;

extern "C" int mobi_layernorm(const mobi_layernorm_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->x && a->out, "mobi_layernorm: null argument");
    MOBI_CHECK(a->C % 4 == 0 && a->C <= LN_MAXV * 128, "mobi_layernorm: C=%d must be a multiple of 4 and <= %d", a->C,
               LN_MAXV * 128);
    MOBI_CHECK(a->gamma == nullptr || a->beta != nullptr, "mobi_layernorm: gamma without beta");
    if (a->rows == 0) return 0;
    MOBI_CHECK(a->rows < (1ll << 31) && a->seg < (1ll << 31) && a->seg_stride < (1ll << 31) &&
                   a->seg_offset < (1ll << 31) && a->add_rows_per_vec < (1ll << 31),
               "mobi_layernorm: row counts must fit int32");
    const int seg = a->seg > 0 ? (int)a->seg : (int)a->rows;
    const int seg_stride = a->seg > 0 ? (int)a->seg_stride : (int)a->rows;
    const int rpv = a->add_rows_per_vec > 0 ? (int)a->add_rows_per_vec : 1;
    const int warps = 8;
#define LN_LAUNCH(MV)                                                                                                  \
    MOBI_CUDA(launch_pdl(layernorm_kernel<MV>, dim3((unsigned)((a->rows + warps - 1) / warps)), dim3(warps * 32), (size_t)0, \
                         stream, reinterpret_cast<float*>(a->x), a->gamma, a->beta,                                    \
                         reinterpret_cast<__nv_bfloat16*>(a->out), a->add_vec, (int)a->rows, a->C, seg, seg_stride,      \
                         (int)a->seg_offset, rpv, a->eps))
    if (a->C <= 384) LN_LAUNCH(3);
    else if (a->C <= 768) LN_LAUNCH(6);
    else LN_LAUNCH(12);
#undef LN_LAUNCH
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_ln_dual(const float* x, const mobi_ln_dual_spec* spec, int32_t batch, int32_t tokens, int32_t C,
                            float eps, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(x && spec, "mobi_ln_dual: null argument");
    MOBI_CHECK(C % 4 == 0 && C <= LN_MAXV * 128, "mobi_ln_dual: C=%d must be a multiple of 4 and <= %d", C, LN_MAXV * 128);
    MOBI_CHECK(!spec->pair || batch % 2 == 0, "mobi_ln_dual: paired mode needs an even batch (got %d)", batch);
    DualParams d;
    if (fill_dual(d, spec)) return 1;
    const long long rows = (long long)batch * tokens;
    if (rows == 0) return 0;
    MOBI_CHECK(rows < (1ll << 31), "mobi_ln_dual: too many rows");
    const int warps = 8;
    const unsigned blocks = (unsigned)((rows + warps - 1) / warps);
    if (C <= 384) MOBI_CUDA(launch_pdl(ln_dual_kernel<3>, dim3(blocks), dim3(warps * 32), (size_t)(0), stream, x, d, (int)rows, C, tokens, eps));
    else if (C <= 768) MOBI_CUDA(launch_pdl(ln_dual_kernel<6>, dim3(blocks), dim3(warps * 32), (size_t)(0), stream, x, d, (int)rows, C, tokens, eps));
    else MOBI_CUDA(launch_pdl(ln_dual_kernel<12>, dim3(blocks), dim3(warps * 32), (size_t)(0), stream, x, d, (int)rows, C, tokens, eps));
    MOBI_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int mobi_ln_adapter(const mobi_ln_adapter_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    MOBI_CHECK(a && a->x && a->gamma && a->beta && a->Ug && a->sb && a->Z && a->zb, "mobi_ln_adapter: null argument");
    MOBI_CHECK(a->C % 4 == 0 && a->C <= LN_MAXV * 128, "mobi_ln_adapter: C=%d must be a multiple of 4 and <= %d", a->C,
               LN_MAXV * 128);
    MOBI_CHECK(a->batch > 0 && a->tokens > 0, "mobi_ln_adapter: empty problem");
    MOBI_CHECK(!a->next.pair || a->batch % 2 == 0, "mobi_ln_adapter: paired mode needs an even batch");
    AdapterParams p;
    p.x = a->x;
    p.add_vec = a->add_vec;
    p.gamma = a->gamma;
    p.beta = a->beta;
    p.Ug = a->Ug;
    p.sb = a->sb;
    p.Z = a->Z;
    p.zb = a->zb;
    p.T = a->tokens;
    p.C = a->C;
    p.eps = a->eps;
    DualParams d;
    if (fill_dual(d, &a->next)) return 1;
    // rows per CTA: amortise the 128*C-byte table load, but keep >= ~2 waves of CTAs
    int rpc = 128;
    while (rpc > 16 && (long long)a->batch * ((a->tokens + rpc - 1) / rpc) < 2 * sm_count()) rpc >>= 1;
    p.rows_per_cta = rpc;
    const size_t smem = (size_t)(2 * AD_H + 1) * a->C * sizeof(float);
    static bool configured = false;
    if (!configured) {
        MOBI_CUDA(cudaFuncSetAttribute(ln_adapter_kernel<3, AD_R3, AD_B3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MOBI_CUDA(cudaFuncSetAttribute(ln_adapter_kernel<6, AD_R6, AD_B6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MOBI_CUDA(cudaFuncSetAttribute(ln_adapter_kernel<12, AD_R12, AD_B12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    MOBI_CHECK(smem <= 200 * 1024, "mobi_ln_adapter: C=%d needs %zu bytes of shared memory", a->C, smem);
    dim3 grid((a->tokens + rpc - 1) / rpc, a->batch);
    if (a->C <= 384) MOBI_CUDA(launch_pdl(ln_adapter_kernel<3, AD_R3, AD_B3>, dim3(grid), dim3(AD_WARPS * 32), (size_t)(smem), stream, p, d));
    else if (a->C <= 768) MOBI_CUDA(launch_pdl(ln_adapter_kernel<6, AD_R6, AD_B6>, dim3(grid), dim3(AD_WARPS * 32), (size_t)(smem), stream, p, d));
    else MOBI_CUDA(launch_pdl(ln_adapter_kernel<12, AD_R12, AD_B12>, dim3(grid), dim3(AD_WARPS * 32), (size_t)(smem), stream, p, d));
    MOBI_CUDA(cudaGetLastError());
    return 0;
}
