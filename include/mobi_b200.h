/*
 * mobi_b200 — C ABI of the B200-native (sm_100a) kernels behind MObI's denoising hot path.
 *
 * The reference (alexbuburuzan/MObI) has no FFI of its own on this path: every operator is a stock
 * torch.nn module created through `instantiate_from_config` (ldm/util.py:76-91).  This header is the
 * boundary a maintainer binds instead: the Python drop-in classes in `mobi_b200/` (same class names,
 * constructor kwargs and state-dict keys as the reference) call these entry points through ctypes with
 * raw device pointers.  Each entry cites the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); nothing is allocated or freed here;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises, and is
 *     CUDA-graph capturable;
 *   - return value 0 = ok, nonzero = error; `mobi_last_error()` returns a thread-local message;
 *   - activations are channels-last: images are [N, H, W, C], token matrices are [rows, C];
 *   - "bf16" is __nv_bfloat16, "f32" is float.  dtype codes: 0 = bf16, 1 = f32.
 */
#ifndef MOBI_B200_H
#define MOBI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOBI_DTYPE_BF16 0
#define MOBI_DTYPE_F32 1

/* Epilogue layouts of mobi_gemm (what is done with each fp32 accumulator tile). */
#define MOBI_EPI_PLAIN 0   /* out[m, n]                                     (row stride ldo)            */
#define MOBI_EPI_GEGLU 1   /* columns come in 16-groups [8 value | 8 gate]; out[m, n/2] = v * gelu(g)  */
#define MOBI_EPI_HEADS 2   /* out[((m / tokens) * heads + n / d) , m % tokens, n % d]   (q, k)          */
#define MOBI_EPI_HEADS_T 3 /* out[((m / tokens) * heads + n / d) , n % d, m % tokens]   (v transposed)  */
#define MOBI_EPI_QKV 4     /* n / (heads*d) selects q (HEADS -> out), k (HEADS -> out2), v (HEADS_T -> out3) */
#define MOBI_EPI_KV 5      /* n / (heads*d) selects k (HEADS -> out), v (HEADS_T -> out2)                     */
#define MOBI_EPI_QKV_ROW 7 /* like QKV but v is written like k: [BH, tokens, d] (mobi_attention v_rowmajor = 1)  */
#define MOBI_EPI_KV_ROW 8  /* like KV  but v is written like k                                                 */
#define MOBI_EPI_GEGLU2 6  /* columns come in (value, gate) PAIRS: out[m, n/2] = v * gelu(g); the layout the
                              persistent kernel's vector epilogue handles without cross-lane traffic          */

const char* mobi_last_error(void);
int mobi_version(void);

/*
 * C[M,N] = A[M,K] . B[N,K]^T on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM, TMA-staged tiles),
 * followed by a fused epilogue:  v = acc + bias[n] + row_bias[m / rows_per_group, n] + residual[m, n].
 *
 * Replaces torch.nn.Linear / 1x1 nn.Conv2d everywhere on the path: CrossAttention.to_q/to_k/to_v/to_out
 * (ldm/modules/attention.py:162-169), GEGLU.proj + FeedForward.net[2] (attention.py:38-62), the adapter
 * connectors (attention.py:218-223), SpatialTransformer.proj_in/proj_out (attention.py:285-300), ResBlock
 * skip_connection (openaimodel.py:241), emb_layers / time_embed (openaimodel.py:204-210, 627-631), and the
 * VAE nin_shortcut / AttnBlock q,k,v,proj_out (model.py:115-119, 156-175).
 *
 * With `conv = 1` the A operand is an NHWC image and the K loop walks the KHxKW filter taps with shifted,
 * zero-filled 4-D TMA boxes (implicit GEMM): replaces nn.Conv2d 3x3 / (1,5), stride 1, "same" zero padding
 * (openaimodel.py:192, 216, 115; model.py:45, 96, 106, 384-401, 559-578).  B is then [N, KH*KW*C] with K
 * ordered (kh, kw, c).
 */
typedef struct {
    const void* A;         /* bf16 [M, K] row-major (lda), or NHWC image when conv=1 */
    const void* B;         /* bf16 [N, K] row-major (ldb) */
    void* out;             /* see epilogue */
    void* out2;            /* MOBI_EPI_QKV: k */
    void* out3;            /* MOBI_EPI_QKV: v (transposed) */
    const float* bias;     /* f32 [N] or NULL */
    const float* row_bias; /* f32 [ceil(M / rows_per_group), N] or NULL */
    const void* residual;  /* [M, N] (ldo) of res_dtype or NULL; may alias out */
    int64_t M, N, K;
    int64_t lda, ldb, ldo; /* row strides in elements */
    int64_t rows_per_group;
    int64_t ld_row_bias;   /* row stride of row_bias in elements (0 = N) */
    int32_t out_dtype, res_dtype;
    int32_t epilogue;
    int32_t act;           /* applied after bias (PLAIN only): 0 = none, 1 = SiLU (time_embed, openaimodel.py:627-631;
                              BBoxEmbedder, encoders/modules.py:195-201), 2 = exact-erf GELU (xf.py MLP, xf.py:49-60) */
    int32_t heads, head_dim, tokens;
    /* PLAIN only: output (and residual) row of GEMM row m is (m / out_seg) * out_seg_stride + out_seg_offset +
     * m % out_seg; out_seg = 0 means identity.  Writes the camera-only / lidar-only token rows of the
     * interleaved batch in place (attention.py:246-263). */
    int64_t out_seg, out_seg_stride, out_seg_offset;
    /* implicit conv */
    int32_t conv;
    int32_t n_img, H, W, C, KH, KW, pad_h, pad_w;
    int32_t tile_n; /* 0 = choose */
    int32_t kernel; /* 0 = choose (persistent kernel when its epilogue supports the problem), 1 = one-tile kernel */
    /* Batched GEMM (PLAIN epilogue, persistent kernel only): `batch` independent problems of the same M, N, K whose
     * operands / outputs are `*_batch_stride` ELEMENTS apart (A, B, out, residual; bias is shared).  batch <= 1 means
     * one problem.  The per-(row, head) products of the attention backward (attention.py:179-192 under autograd)
     * run as one launch per batch row with batch = heads. */
    int32_t batch;
    int64_t a_batch_stride, b_batch_stride, out_batch_stride;
    /* CTA pairs (tcgen05 cta_group::2, 256-row tiles): 0 = choose, 1 = force on (persistent kernel only), -1 = off,
     * 2 = clusters of two pairs on neighbouring n-tiles that fetch their common A tiles in halves and multicast them
     * (PLAIN epilogue, tile_n >= 128, even number of n-tiles, no batch; a tuning arm: never chosen automatically),
     * 3 = pairs on 256 x 320 tiles (tile_n = 160, two N = 160 MMAs per k-step on one A tile, three rotating TMEM
     * accumulator slots; no batch): chosen automatically for convolutions with N % 320 == 0 and a full wave of tiles. */
    int32_t pair;
    /* MN-major ("transposed") operands, persistent kernel only, no conv / CTA pairs: with a_mn_major, A is given as the
     * row-major [K, M] matrix (lda = its row stride >= M) and read as A^T by the tensor core (tcgen05 instruction-descriptor
     * bit 15, MN-major shared-memory descriptors); with b_mn_major, B is given as row-major [K, N] (ldb >= N; tile_n a
     * multiple of 64).  out = A_given^T . B_given is exactly the weight gradient dY^T X of a Linear over token rows and
     * the dK / dV / dQ products of the attention backward, so none of them needs a transposed copy.  With both flags K
     * may be any positive count (rows beyond K are zero-filled by TMA). */
    int32_t a_mn_major, b_mn_major;
    /* out (f32, PLAIN epilogue, no bias / activation / residual) is accumulated with fp32 atomic adds instead of stored:
     * with batch > 1 and out_batch_stride = 0 the batch entries are K-slices of ONE product (split-K), e.g. a weight
     * gradient over 16,384 token rows whose 6 output tiles would otherwise occupy 6 of 148 SMs. */
    int32_t atomic_out;
    /* Two-level batch: with batch_inner > 0 the batch index z splits into (z % batch_inner, z / batch_inner); the inner
     * level uses the *_batch_stride fields, the outer level these (elements).  Heads inside batch rows: head h of batch row
     * b of a token-major [B * T, H * D] matrix sits at b * T * ld + h * D, which no single stride expresses. */
    int32_t batch_inner;
    int64_t a_batch2_stride, b_batch2_stride, out_batch2_stride;
    /* Optional (PLAIN epilogue, f32 output, persistent kernel, no batch / row segments): per-column sums and sums of
     * squares of the FINAL output values (after bias, row bias, activation and residual) over every group of 32
     * consecutive output rows, as f32 [2][ceil(M / 32)][N] (plane 0 = sums, plane 1 = sums of squares).  They are the
     * statistics the GroupNorm that consumes this output needs (openaimodel.py:255-275, attention.py:302-313), computed
     * while the tile is still in registers: mobi_groupnorm then reads its input once instead of twice.  Per-COLUMN sums
     * are consumer-agnostic: the same tensor feeds GroupNorms with different group widths (as h, and later as a skip
     * connection inside a wider concatenation, openaimodel.py:892). */
    float* colstats;
} mobi_gemm_args;

int mobi_gemm(const mobi_gemm_args* args, void* stream);
/* What mobi_gemm would do with these arguments, without touching the device: the tile width it picks (64 / 128 / 160 / 256),
 * the CTA grouping (the values of `pair` above; 0 = one CTA per tile) and whether the persistent kernel takes the problem.
 * Same validation and error behaviour as mobi_gemm; pointers are only checked for null / alignment, never dereferenced. */
int mobi_gemm_plan(const mobi_gemm_args* args, int32_t* tile_n, int32_t* pair, int32_t* persistent);

/*
 * Flash-style fused attention: O = softmax(Q K^T) V per (batch*head), scores never leave the SM.
 * q, k: bf16 [BH, Tq|Tk, d]; vt: bf16 [BH, d, Tk] (transposed V); out: bf16 [B, Tq, heads*d] token-major.
 * The softmax scale (d^-0.5, attention.py:159,181; C^-0.5 in the VAE AttnBlock, model.py:191) times log2(e)
 * must already be folded into q.  Replaces CrossAttention.forward lines 179-193 and AttnBlock.forward
 * lines 184-198.  Cross-modal attention (attention.py:246-261) is the same call
 * with q built from one modality's rows and k/vt from the partner rows.
 */
typedef struct {
    const void* q;
    const void* k;
    const void* vt;
    void* out;
    int32_t batch, heads, head_dim, tq, tk;
    int64_t ld_out; /* row stride of out in elements (>= heads*head_dim) */
    int32_t kernel; /* 0 = choose (two-tile kernel for head_dim <= 128), 1 = force the one-tile kernel,
                       2 = two-tile kernel with 128-key blocks and one CTA per SM (head_dim <= 64) */
    int32_t v_rowmajor; /* 1: `vt` holds V as [BH, Tk, d] (same layout as k; head_dim <= 128): consumed as an MN-major
                           tcgen05 operand, no transposed copy needed.  0: `vt` is V^T [BH, d, Tk]. */
    float* lse;         /* optional f32 [BH, tq] (v_rowmajor = 1 only): log2-domain log-sum-exp of every score row,
                           rowmax + log2(rowsum); the training step keeps it so that the backward needs no statistics pass */
} mobi_attn_args;

int mobi_attention(const mobi_attn_args* args, void* stream);

/*
 * GroupNorm (+ optional SiLU) over an NHWC f32 or bf16 image, fp32 statistics (GroupNorm32,
 * ldm/modules/diffusionmodules/util.py:214-216; eps 1e-5 in ResBlocks, 1e-6 in SpatialTransformer.norm
 * and the VAE, attention.py:77-78, model.py:38-39), fused with the following SiLU (openaimodel.py:190,
 * 213; model.py:33-35).  The input may be the channel concatenation of two tensors (x1 | x2), which is how
 * UNetModel.forward feeds skip connections (openaimodel.py:892).  Output bf16 NHWC [N, HW, C1+C2].
 * `partials` is caller-provided scratch of mobi_groupnorm_scratch_bytes() bytes.
 */
typedef struct {
    const void* x1;
    const void* x2; /* NULL when there is no concatenation */
    const float* gamma;
    const float* beta;
    void* out;        /* [N, HW, C] of out_dtype */
    void* out_concat; /* optional: the raw (un-normalised) concatenation as bf16 [N, HW, C] (NULL to skip) */
    float* partials;
    int32_t n_img, hw, c1, c2, groups;
    int32_t in_dtype;
    int32_t silu;
    float eps;
    int32_t out_dtype; /* dtype of `out`: MOBI_DTYPE_BF16 (default, a GEMM/conv operand) or MOBI_DTYPE_F32 */
    int32_t force_two_pass; /* 1: always run the statistics + apply kernel pair (default 0: single-pass cluster kernel
                               whenever one cluster's slab set fits L2) */
    /* Optional: the per-column statistics mobi_gemm left for x1 / x2 (mobi_gemm_args.colstats; f32 [2][n_img * hw / 32][c];
     * hw must be a multiple of 32).  When given for every input, no statistics pass runs: a small kernel folds the column
     * sums into the (image, group) sums and the apply kernel streams the input once (4 B read + 2 B written per element). */
    const float* colstats1;
    const float* colstats2;
} mobi_groupnorm_args;

int64_t mobi_groupnorm_scratch_bytes(int32_t n_img, int32_t hw, int32_t c, int32_t groups);
/* Kernels mobi_groupnorm launches for this shape: 1 (single-pass cluster kernel) or 2 (statistics + apply). */
int32_t mobi_groupnorm_launches(int32_t hw, int32_t c, int32_t groups, int32_t in_dtype, int32_t force_two_pass);
int mobi_groupnorm(const mobi_groupnorm_args* args, void* stream);

/*
 * LayerNorm over the last dim of an f32 token matrix -> bf16 (nn.LayerNorm, eps 1e-5, attention.py:213-223).
 * Row gather: output row i reads input row  (i / seg) * seg_stride + seg_offset + i % seg , which selects the
 * camera (even) or lidar (odd) batch rows of the interleaved batch (attention.py:246-247) without a copy.
 * gamma == NULL means "cast only" (the un-normalised partner tokens used as cross-modal context).
 * add_vec: optional f32 [rows / add_rows_per_vec, C] added to the input row before normalising, and written
 * back to x (x += vec) — the broadcast form of the single-key attn2 (attention.py:235).
 */
typedef struct {
    void* x; /* f32 [*, C] (read; written when add_vec != NULL) */
    const float* gamma;
    const float* beta;
    void* out; /* bf16 [rows, C] */
    const float* add_vec;
    int64_t rows;
    int32_t C;
    int64_t seg, seg_stride, seg_offset;
    int64_t add_rows_per_vec;
    float eps;
} mobi_layernorm_args;

int mobi_layernorm(const mobi_layernorm_args* args, void* stream);

/*
 * One pass over ALL token rows of the interleaved batch x f32 [batch, tokens, C]: with pair = 1, even batch rows
 * (camera) take slot 0 and odd batch rows (lidar) slot 1 (attention.py:246-247); each slot has its own action
 * (mode 0 = skip, 1 = LayerNorm with its own gamma/beta, 2 = plain cast) and its own COMPACT bf16 output
 * [batch/2, tokens, C].  With pair = 0 every row takes slot 0 and the output is [batch, tokens, C].
 * Produces, in one launch, the normalised queries of one modality and the un-normalised context tokens of the other
 * (attention.py:249-261).
 */
typedef struct {
    const float* gamma[2];
    const float* beta[2];
    void* out[2];
    int32_t mode[2];
    int32_t pair;
} mobi_ln_dual_spec;
int mobi_ln_dual(const float* x, const mobi_ln_dual_spec* spec, int32_t batch, int32_t tokens, int32_t C, float eps,
                 void* stream);

/*
 * The bbox / reference-image adapter (cond_adapter_norm -> cond_adapter_attn -> cond_adapter_connector,
 * attention.py:237-243) in folded form, fused with the pending attn2 vector add (attention.py:235) before it and
 * the LayerNorm(s) after it.  Per token row of batch row b:
 *     x += add_vec[b]                                  (optional)
 *     d = (x - mean) * rstd                            (cond_adapter_norm statistics, eps)
 *     s_j = <d, Ug[b, j]> + sb[b, j]     j = key*8 + head, 2 keys x 8 head slots (unused heads zero)
 *     p = softmax over the 2 keys of each head ;  x += sum_j p_j Z[b, j] + zb ;  x written back
 *     then `next` is applied to the updated row exactly like mobi_ln_dual.
 * Ug = gamma * (W_q^T k * scale), sb = <beta, W_q^T k * scale>, Z = (W_connector W_out) v: all f32, built once per
 * sampling run from the context.
 */
typedef struct {
    float* x;             /* f32 [batch, tokens, C], updated in place */
    const float* add_vec; /* f32 [batch, C] or NULL */
    const float* gamma;
    const float* beta;
    const float* Ug; /* f32 [batch, 16, C] */
    const float* sb; /* f32 [batch, 16] */
    const float* Z;  /* f32 [batch, 16, C] */
    const float* zb; /* f32 [C] */
    int32_t batch, tokens, C;
    float eps;
    mobi_ln_dual_spec next;
} mobi_ln_adapter_args;
int mobi_ln_adapter(const mobi_ln_adapter_args* args, void* stream);

/* timestep_embedding (ldm/modules/diffusionmodules/util.py:151-171): out bf16 [n, dim] = [cos | sin]. */
int mobi_timestep_embedding(const int64_t* t, void* out_bf16, int32_t n, int32_t dim, float max_period,
                            void* stream);

/* Fourier features of the BBoxEmbedder (ldm/modules/encoders/modules.py:203-206, 215-252): x f32 [rows, dims] ->
 * out bf16 [rows, dims * (1 + 2 * num_freqs)] = [x, sin(x f_0), cos(x f_0), sin(x f_1), ...], f_i = 2^i, zero-padded to
 * ld_out columns. */
int mobi_fourier_embed(const float* x, void* out_bf16, int64_t rows, int32_t dims, int32_t num_freqs, int64_t ld_out,
                       void* stream);

/* y = silu(x) elementwise, f32 or bf16 in -> bf16 out (emb_layers[0], openaimodel.py:204). */
int mobi_silu(const void* x, int32_t in_dtype, void* out_bf16, int64_t n, void* stream);

/* Layout changes at the module boundary: NCHW f32 <-> NHWC (f32 or bf16). */
int mobi_nchw_to_nhwc(const float* x, void* out, int32_t out_dtype, int32_t n, int32_t c, int32_t hw,
                      void* stream);
int mobi_nhwc_to_nchw(const void* x, int32_t in_dtype, float* out, int32_t n, int32_t c, int32_t hw,
                      void* stream);

/* Nearest-neighbour x2 upsampling of an NHWC image (F.interpolate(scale_factor=2, mode="nearest"),
 * openaimodel.py:116, model.py:53).  in/out dtype f32 or bf16 (converted on the fly). */
int mobi_upsample_nearest2x(const void* x, int32_t in_dtype, void* out, int32_t out_dtype, int32_t n, int32_t h,
                            int32_t w, int32_t c, void* stream);

/* Explicit im2col (bf16 NHWC -> bf16 [N*Ho*Wo, Kpad]) for the few convolutions the implicit path does not
 * take: stride 2 (Downsample, openaimodel.py:151-153; VAE Downsample with pad (0,1,0,1), model.py:72-76)
 * and tiny channel counts (first conv 9->320, openaimodel.py:666).  K order (kh, kw, c), zero padded to kpad. */
typedef struct {
    const void* x; /* NHWC in_dtype */
    void* out;     /* bf16 [n*ho*wo, kpad] */
    int32_t in_dtype;
    int32_t n, h, w, c, kh, kw, stride, pad_top, pad_left, ho, wo, kpad;
} mobi_im2col_args;
int mobi_im2col(const mobi_im2col_args* args, void* stream);

/*
 * Folded 2-key cross attention (cond_adapter_attn + connector, attention.py:237-243): per token
 *   s[h, j] = xn . U[b, h*2+j, :]      (U already holds W_q^T k * scale)
 *   p = softmax_j(s) ;  x += sum_{h,j} p[h,j] * Z[b, h*2+j, :] + zb
 * xn: bf16 [B*T, C] (LayerNorm output), U: f32 [B, 2*heads, C], Z: f32 [B, 2*heads, C], zb: f32 [C],
 * x: f32 [B*T, C] updated in place.
 */
typedef struct {
    const void* xn;
    const float* U;
    const float* Z;
    const float* zb;
    float* x;
    int32_t batch, tokens, C, heads, keys;
} mobi_ctx_attn_args;
int mobi_ctx_attention(const mobi_ctx_attn_args* args, void* stream);

/*
 * One sampler update (DDIMSampler.p_sample_ddim, ldm/models/diffusion/ddim.py:184-212 and
 * PLMSSampler.p_sample_plms, plms.py:200-237), fused:
 *   e      = e_uncond + scale * (e_cond - e_uncond)                       (CFG, ddim.py:184)
 *   e'     = c0*e + c1*old1 + c2*old2 + c3*old3                           (PLMS multistep, plms.py:221-235; DDIM: c0=1)
 *   pred   = (x - sqrt(1-a_t) * e') / sqrt(a_t)                           (ddim.py:201-204)
 *   x_prev = sqrt(a_prev) * pred + sqrt(1 - a_prev - sigma^2) * e' + sigma * noise * temperature  (ddim.py:208-212)
 * and the next UNet input  [x_prev | inpaint_image | inpaint_mask]  (ddim.py:172) is assembled for both CFG
 * halves.  All NCHW f32.  eps holds [uncond ; cond] rows when cfg != 0.
 */
typedef struct {
    const float* eps;   /* [(cfg?2:1)*B, 4, H, W] */
    const float* x;     /* [B, 4, H, W] */
    const float* noise; /* [B, 4, H, W] or NULL */
    const float* old1;
    const float* old2;
    const float* old3; /* previous combined eps (PLMS) or NULL */
    float* e_out;      /* optional: CFG-combined eps e (before multistep) [B,4,H,W] */
    float* x_prev;     /* [B, 4, H, W] */
    float* pred_x0;    /* [B, 4, H, W] */
    int64_t n;         /* B*4*H*W */
    int32_t cfg;
    float scale, c0, c1, c2, c3;
    float sqrt_one_minus_at, sqrt_at, sqrt_a_prev, dir_coef, sigma_temp;
} mobi_sampler_args;
int mobi_sampler_update(const mobi_sampler_args* args, void* stream);

/* x_in[(cfg?2:1)*B, 9, H, W] = cat([x, inpaint_image, inpaint_mask], 1) repeated for the CFG halves
 * (ddim.py:172,180).  Optional known-latent blend first: x = x0_noised*mask + (1-mask)*x (ddim.py:145-148),
 * where x0_noised = sqrt_ac * x0 + sqrt_1mac * noise (q_sample, ddpm.py:284-287). */
typedef struct {
    float* x;               /* [B,4,H,W]; updated in place when blend_mask != NULL */
    const float* inpaint_image; /* [B,4,H,W]; or the whole `rest` tensor [B,rest_c,H,W] when inpaint_mask is NULL */
    const float* inpaint_mask;  /* [B,1,H,W] or NULL (ddim.py:173-176: kwargs['rest']) */
    const float* blend_mask;    /* [B, blend_c, H, W], blend_c = 1 (broadcast over channels) or 4; NULL = no blend */
    const float* blend_x0;
    const float* blend_noise;
    float* x_in; /* [(cfg?2:1)*B, 9, H, W] */
    int32_t B, hw, cfg, blend_c, rest_c;
    float sqrt_ac, sqrt_1mac;
} mobi_assemble_args;
int mobi_assemble_input(const mobi_assemble_args* args, void* stream);

/* Row softmax p = 2^(s - rowmax) / rowsum of GEMM-produced scores (f32 [rows, cols], row stride ld_s) -> bf16
 * (row stride ld_p).  The VAE AttnBlock (model.py:184-195) has ONE head of d = 512: its QK^T and PV products run as
 * two mobi_gemm calls per image around this kernel; the C^-0.5 * log2(e) scale is folded into the q projection. */
int mobi_softmax_rows(const float* s, void* p_bf16, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p,
                      void* stream);

/* out[i] = a[i] + b[i] (f32), used for residual joins that have no GEMM to fuse into. */
int mobi_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream);
/* out = x * s (f32 -> f32), e.g. z / scale_factor before decode (ddpm.py:846-849). */
int mobi_scale_f32(const float* x, float s, float* out, int64_t n, void* stream);
/* f32 -> bf16 cast */
int mobi_cast_bf16(const float* x, void* out, int64_t n, void* stream);
/* dst[i] = bf16(src[i] * seg_scale[j]) for i in [seg_start[j], seg_start[j + 1]) (sorted starts, multiples of 4, seg_start[0]
 * = 0; the last segment runs to n): the whole flat f32 master-weight buffer of the training step becomes the bf16 operand
 * packs of the next forward / backward in ONE launch, with the attention scale (and log2 e) folded into the to_q
 * matrices (attention.py:171-177; what torch autocast / .to(bf16) would do per tensor). */
int mobi_cast_bf16_segments(const float* src, void* dst, int64_t n, const int64_t* seg_start, const float* seg_scale,
                            int32_t nseg, void* stream);
/* x[i] *= seg_scale[j] over the same kind of segment table, in place, skipping the blocks whose scale is 1: the chain rule
 * through the scale folded into the packed to_q matrices (gradients w.r.t. W_q * scale -> w.r.t. W_q), one launch. */
int mobi_scale_segments(float* x, int64_t n, const int64_t* seg_start, const float* seg_scale, int32_t nseg, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step (config 5): LatentDiffusion.forward / p_losses (ldm/models/diffusion/ddpm.py:1040-1058, 1177-1217) and
 * what torch.autograd executes behind `loss.backward()` for the UNet (trainable = cond_adapter_* and cross_modal_*
 * parameters, ddpm.py:1686-1698).  The contractions of the backward pass (dgrad, wgrad, the five products of the
 * attention backward) are mobi_gemm calls on transposed / flipped weight packs; the entries below are the rest.
 * ------------------------------------------------------------------------------------------------------------------ */

/* out[b][c, r] = bf16(in[b][r, c]); in f32 or bf16, row strides in elements.  Builds the K-major operands of
 * wgrad (dW = dY^T X contracts over tokens) and of the attention backward (K^T, Q^T, dO^T). */
int mobi_transpose_bf16(const void* in, int32_t in_dtype, void* out_bf16, int64_t batch, int32_t rows, int32_t cols,
                        int64_t ld_in, int64_t in_batch_stride, int64_t ld_out, int64_t out_batch_stride, void* stream);

/* Backward of nn.LayerNorm (attention.py:213-223) for the rows selected by the same segment gather as
 * mobi_layernorm: dx[row] (+)= dLN(x[row], dy[i]); dgamma/dbeta (f32 [C], atomically accumulated) only for the
 * trainable adapter norms.  gamma == NULL means unit gamma. */
typedef struct {
    const float* x;
    const float* gamma;
    const void* dy; /* [rows, C] compact, dy_dtype */
    float* dx;      /* f32, addressed like x */
    float* dgamma;
    float* dbeta;
    int64_t rows;
    int32_t C;
    int64_t seg, seg_stride, seg_offset;
    float eps;
    int32_t dy_dtype;
    int32_t accumulate; /* 1: dx += ... */
} mobi_layernorm_bwd_args;
int mobi_layernorm_bwd(const mobi_layernorm_bwd_args* args, void* stream);

/* Backward of GroupNorm (+SiLU) over an NHWC f32 image that may be the channel concatenation (x1 | x2)
 * (GroupNorm32 + SiLU of ResBlock, openaimodel.py:190, 213, 892; SpatialTransformer.norm, attention.py:285):
 * dx = dGN(dy) + dres, split back into dx1 [N,HW,c1] and dx2 [N,HW,c2].  The norms are frozen: no dgamma. */
typedef struct {
    const float* x1;
    const float* x2;
    const float* gamma;
    const float* beta;
    const void* dy;    /* [N, HW, C] dy_dtype: gradient w.r.t. the (SiLU'd) normalised output */
    const float* dres; /* optional f32 [N, HW, C] added to dx (the identity branch of the residual) */
    float* dx1;
    float* dx2;
    int32_t n_img, hw, c1, c2, groups, silu, dy_dtype;
    float eps;
} mobi_groupnorm_bwd_args;
int mobi_groupnorm_bwd(const mobi_groupnorm_bwd_args* args, void* stream);

/* GEGLU (attention.py:38-45) with exact erf GELU on (value, gate) column PAIRS: g bf16 [rows, 2*features] ->
 * out bf16 [rows, features] = value * gelu(gate); backward: dg from g and dh. */
int mobi_geglu(const void* g, void* out, int64_t rows, int64_t features, void* stream);
int mobi_geglu_bwd(const void* g, const void* dh, void* dg, int64_t rows, int64_t features, void* stream);

/* Softmax backward of CrossAttention.forward (attention.py:181-190) from materialised f32 tiles S = q' k^T (log2
 * domain) and dP = dO V^T, both [batch, tq, tk]:  P = softmax2(S), dS = dscale * P * (dP - rowsum(P * dP)).
 * Writes dS [tq, tk] and the transposed dS^T, P^T [tk, tq] (bf16): the A operands of dQ = dS k, dK = dS^T q',
 * dV = P^T dO.  stats: scratch f32 [batch * tq * 3]. */
typedef struct {
    const float* S;
    const float* dP;
    void* dS;
    void* dSt;
    void* Pt;
    float* stats;
    int32_t batch, tq, tk;
    float dscale;
} mobi_attn_softmax_bwd_args;
int mobi_attn_softmax_bwd(const mobi_attn_softmax_bwd_args* args, void* stream);
/* Same, with the row statistics taken from the forward instead of a pass over S and dP: `lse` f32 [batch, tq] from
 * mobi_attention, Delta = rowsum(dO * O) computed here from the token-major bf16 o / d_o (row stride ld, head columns
 * [h*d, (h+1)*d), batch entry = head h of one batch row).  args->stats is still the scratch it fills. */
int mobi_attn_softmax_bwd_lse(const mobi_attn_softmax_bwd_args* args, const float* lse, const void* o, const void* d_o,
                              int64_t ld, int32_t head_dim, void* stream);

/* Fused form of the two score products + mobi_attn_softmax_bwd for head_dim <= 128 and tokens % 128 == 0: S = q' k^T
 * and dP = dO v^T are recomputed tile by tile on tcgen05 (two TMEM accumulators per 128 x 128 tile) in a statistics pass
 * (rowmax, 1 / rowsum, Delta -> stats) and a main pass that writes dS, dS^T, P^T (bf16 [heads, tokens, tokens]) directly:
 * the f32 T x T tiles never reach HBM.  q, k, v: bf16 [batch_rows * heads, tokens, head_dim] (q' carries
 * scale * log2(e)); d_o: bf16 token-major [batch_rows * tokens, ld_do], head h = columns [h * head_dim, (h + 1) * head_dim).
 * stats: f32 [batch_rows * heads, tokens, 3]; dS, dSt, Pt: bf16 [batch_rows * heads, tokens, tokens].  With dSt == NULL
 * the kernel writes P ROW-MAJOR into Pt and no transposed tiles at all: dK = dS^T q' and dV = P^T dO then read dS / P as
 * MN-major GEMM operands (mobi_gemm a_mn_major).
 * stats_only = 1 runs the first pass only; batch_rows = 0 means 1. */
typedef struct {
    const void* q;
    const void* k;
    const void* v;
    const void* d_o;
    float* stats;
    void* dS;
    void* dSt;
    void* Pt;
    int32_t heads, tokens, head_dim, stats_only;
    int64_t ld_do;
    float dscale;
    int32_t batch_rows;
} mobi_attn_bwd_tiles_args;
int mobi_attn_bwd_tiles(const mobi_attn_bwd_tiles_args* args, void* stream);

/* Flash-style attention backward (attention.py:179-192 under autograd) for head_dim <= 128, tokens % 128 == 0: given the
 * row statistics of mobi_attn_bwd_tiles (stats_only = 1) it recomputes S = q' k^T and dP = dO v^T per 128 x 128 tile on
 * tcgen05, forms P and dS in registers, stages them as bf16 in shared memory and accumulates
 *   dQ' = dS k (one kernel, CTA = 128 query rows)   and   dK = dS^T q', dV = P^T dO (second kernel, CTA = 128 key rows)
 * in TMEM: no T x T tile ever reaches HBM.  q, k, v: bf16 [batch_rows * heads, tokens, head_dim]; d_o and the outputs are
 * token-major [batch_rows * tokens, ld] matrices whose head h occupies columns [h * head_dim, (h + 1) * head_dim) of the
 * given base pointers (which may point into wider rows: QKV gradient buffers). */
typedef struct {
    const void* q;
    const void* k;
    const void* v;
    const void* d_o;
    const float* stats;
    void* dq;
    void* dk;
    void* dv;
    int64_t ld_do, ld_dq, ld_dk, ld_dv;
    int32_t heads, tokens, head_dim, batch_rows;
    float dscale;
} mobi_attn_bwd_flash_args;
int mobi_attn_bwd_flash(const mobi_attn_bwd_flash_args* args, void* stream);

/* cond_adapter_attn (attention.py:237-243, CrossAttention with `keys` <= 4 context tokens) on projected queries, for
 * the training step where to_q/to_k/to_v are trainable and cannot be folded:
 *   forward : o = softmax_j(<q, k_j>) v_j per head                                    (backward = 0)
 *   backward: dq (bf16), dk / dv (f32 [batch, keys, C], atomically accumulated) from d_o (backward = 1)
 * q, o, d_o, dq: bf16 [batch*tokens, C]; k, v: f32 [batch, keys, C]; the softmax scale is folded into q. */
typedef struct {
    const void* q;
    const float* k;
    const float* v;
    void* o;
    const void* d_o;
    void* dq;
    float* dk;
    float* dv;
    int32_t batch, tokens, C, heads, keys, backward;
} mobi_ctx_attn_qspace_args;
int mobi_ctx_attn_qspace(const mobi_ctx_attn_qspace_args* args, void* stream);

/* out[g, c] += sum over the rows r of group g (r / rows_per_group) of x[r, c]: bias gradients (rows_per_group = rows)
 * and per-batch-row sums.  out is f32 [rows / rows_per_group, cols], accumulated atomically. */
int mobi_colsum(const void* x, int32_t dtype, int64_t rows, int32_t cols, int64_t ld, int64_t rows_per_group, float* out,
                void* stream);
/* out[n, k] += sum_m A[m, n] * B[m, k], all f32, tiny m (the context-token side of the adapter: to_k / to_v wgrad). */
int mobi_wgrad_small(const float* A, const float* B, float* out, int32_t m, int32_t n, int32_t k, int64_t lda, int64_t ldb,
                     int64_t ldo, void* stream);
/* dst[segment row r] += src[r] (src compact [rows, C] f32/bf16): joins camera-only / lidar-only gradients
 * (attention.py:246-261) into the interleaved residual-stream gradient. */
int mobi_scatter_add_rows(const void* src, int32_t src_dtype, float* dst, int64_t rows, int32_t C, int64_t seg,
                          int64_t seg_stride, int64_t seg_offset, void* stream);
/* Adjoint of the stride-2 conv3x3 of Downsample (openaimodel.py:151-153): dy [n,h,w,c] -> bf16 [n,2h,2w,c] with dy at
 * the even positions; a stride-1 conv with the flipped filter then gives dx. */
int mobi_zero_insert2x(const void* dy, int32_t in_dtype, void* z_bf16, int32_t n, int32_t h, int32_t w, int32_t c,
                       void* stream);
/* Adjoint of nearest x2 upsampling (openaimodel.py:116): d f32 [n,2h,2w,c] -> out f32 [n,h,w,c]. */
int mobi_sum2x2(const float* d, float* out, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
/* q_sample on the first c_noised channels, the rest copied (ddpm.py:284-287, 1178-1182).  NCHW f32, t int64 [batch]. */
int mobi_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, const int64_t* t,
                  float* out, int32_t batch, int32_t c_total, int32_t c_noised, int32_t hw, void* stream);
/* loss_sum += sum (pred - target)^2 ; grad = grad_scale * (pred - target)  (get_loss 'l2' + mean, ddpm.py:1196-1210). */
int mobi_mse_grad(const float* pred, const float* target, float* grad, float* loss_sum, int64_t n, float grad_scale,
                  void* stream);
/* dx (f32) = dy * silu'(pre): backward of the SiLUs of the trainable BBoxEmbedder MLP (encoders/modules.py:195-201). */
int mobi_silu_bwd(const void* pre, int32_t pre_dtype, const void* dy, int32_t dy_dtype, float* dx, int64_t n, void* stream);
/* torch.optim.AdamW step over one flat f32 buffer (ddpm.py:1655): g is multiplied by grad_scale first. */
int mobi_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, float bias_corr1, float bias_corr2, float grad_scale, void* stream);
/* The same step with torch.optim.AdamW's handling of parameters whose gradient is None (ddpm.py:1655; Lightning's
 * zero_grad sets unused gradients to None): the flat buffer is cut into up to three segments [0, bound1), [bound1, bound2),
 * [bound2, n) (the UNet's adapters | the bbox_embedder | bbox_uncond_vector); a segment whose device flag
 * flags[s] is <= 0 this step is skipped entirely (no weight decay, no moment decay, no step count), every other segment
 * advances its OWN step counter steps[s] (device int32[3], in/out) and uses the bias corrections of that counter.
 * The flags live on the device so that they can ride the gradient all-reduce (a segment is active when any rank gave
 * it a gradient, as under DDP).  state: device scratch of 9 floats.  Two launches, no host synchronisation. */
int mobi_adamw_segments(float* p, const float* g, float* m, float* v, int64_t n, int64_t bound1, int64_t bound2,
                        const float* flags, int32_t* steps, float* state, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float grad_scale, void* stream);

/* Input assembly before the loop / the training step: one modality of
 * LatentDiffusion.encode_all_stages (ldm/models/diffusion/ddpm.py:1010-1033) fused with the lidar alignment and the
 * camera / lidar interleave of LatentDiffusion.get_input (ddpm.py:797-826):
 *   z        = scale * (mean_gt + exp(0.5 * clamp(logvar_gt, -30, 20)) * noise_gt)       (distributions.py:24-37)
 *   z_inp    = same from the inpaint posterior
 *   mask     = nearest resize of the full-resolution keep mask to the latent size          (F.interpolate, ddpm.py:1020)
 *   out[i * row_stride + row_offset, :, y, x] = [z | z_inp | mask][i, :, y - pad, x + left]  (zero outside: F.pad,
 *   a negative pad crops rows), y, x in [0, S).
 * moments: NCHW f32 [n, 8, hs, ws] (AutoencoderKL.encode(...).parameters); noise: [n, 4, hs, ws] or NULL (posterior mode);
 * mask [n, 1, hm, wm]; out [*, 9, S, S].  Camera: left = pad = 0, row_offset 0; lidar: row_offset 1; row_stride 2. */
typedef struct {
    const float* moments_gt;
    const float* noise_gt;
    const float* moments_inpaint;
    const float* noise_inpaint;
    const float* mask;
    float* out;
    int32_t n, hs, ws, hm, wm, S, left, pad, row_stride, row_offset;
    float scale;
} mobi_latent_input_args;
int mobi_assemble_latent_input(const mobi_latent_input_args* args, void* stream);
/* In-place box-corner re-normalisation of get_input (ddpm.py:815-816): bbox f32 [n_points, 3];
 * x = (x * W - left) / S ; y += pad / S. */
int mobi_bbox_renorm(float* bbox, int64_t n_points, int32_t W, int32_t left, int32_t S, int32_t pad, void* stream);

/* ---- Range-view post-processing after the lidar decode --------------------------------------
 * The reference runs this on the host, one sample at a time, in NumPy / cv2 / numba
 * (scripts/inference_test_bench.py:567-629); these entry points keep it in HBM. */
enum {
    MOBI_RANGE_MAP_NONE = 0,
    MOBI_RANGE_MAP_DEPTH_NORM = 1,   /* depth_normalization          ldm/data/utils.py:537-557 */
    MOBI_RANGE_MAP_DEPTH_UNNORM = 2, /* inverse_depth_normalization  ldm/data/utils.py:560-580 */
    MOBI_RANGE_MAP_INT_UNNORM = 3    /* clamp(-0.5 * log(1 - (x + 1) / 2) - 1, -1, 1)  ldm/models/diffusion/ddpm.py:1541 */
};
/* out[b, i] = map(in[b, i]) for f32 [batch, n] (strides in elements, 0 = n; in == out allowed).  The depth maps take the
 * per-sample min_d / max_d [batch] of the object (nuscenes.py:439-442) and alpha (range_object_norm_scale);
 * clamp_input applies torch.clamp(x, -1, 1) first (ddpm.py:1504). */
typedef struct {
    const float* in;
    float* out;
    const float* min_d;
    const float* max_d;
    int64_t n, in_stride, out_stride;
    int32_t batch, mode, clamp_input;
    double alpha; /* a Python float in the reference: 2 * alpha, alpha - 1, 1 - alpha are formed in double, then cast */
} mobi_range_map_args;
int mobi_range_map(const mobi_range_map_args* args, void* stream);

/* LidarConverter.undo_default_transforms batched over samples as postprocess_range_depth_int does
 * (ldm/data/lidar_converter.py:436-485, 230-287; ldm/data/utils.py:471-505), for one or two channels (depth, intensity):
 * the [crop_h, crop_w] crop is resized to [H, width_crop[b]] — F.avg_pool2d when both sizes divide, else cv2's INTER_NEAREST
 * rule — and pasted into a copy of the original [H, W] sweep at column crop_left[b] % W, wrapping the 360 degree seam.
 * map[c] fuses one of the maps above into the crop load (what ddpm.py:1536, 1541 do to the decoded image before);
 * zero_context replaces the original DEPTH by -1 (utils.py:486-487).  crop_left / width_crop: int64 [batch] on the device. */
typedef struct {
    const float* crop[2];
    const float* orig[2];
    float* out[2];
    const float* min_d;
    const float* max_d;
    const int64_t* crop_left;
    const int64_t* width_crop;
    int64_t crop_batch_stride, orig_batch_stride, out_batch_stride; /* elements; 0 = dense */
    int32_t batch, channels, crop_h, crop_w, H, W, zero_context, clamp_input;
    int32_t map[2];
    double alpha;
} mobi_range_undo_args;
int mobi_range_undo_transforms(const mobi_range_undo_args* args, void* stream);

/* LidarConverter.range2pcd (ldm/data/lidar_converter.py:122-176) for [batch, H, W] sweeps at the base size:
 * metres = (depth + 1) / 2 * depth_max; x = cos(yaw) cos(pitch) m, y = -sin(yaw) cos(pitch) m, z = sin(pitch) m; points
 * with depth_min < m < depth_max are kept IN PIXEL ORDER.  points: f32 [batch, H*W, point_stride] (padded; [3] = label
 * when point_stride > 3, [4] = beam index H-1-row when > 4), index (optional): flat pixel of each kept point
 * (the `label = arange` use of inference_test_bench.py:587-588), count: int32 [batch]. */
typedef struct {
    const float* depth;
    const float* pitch;
    const float* yaw;
    const float* label; /* optional f32 [batch, H*W] */
    float* points;
    int32_t* index; /* optional */
    int32_t* count;
    int32_t batch, H, W, point_stride;
    float depth_min, depth_max;
} mobi_range2pcd_args;
int mobi_range2pcd(const mobi_range2pcd_args* args, void* stream);

/* The save_samples sequence of scripts/inference_test_bench.py:580-629 for a batch in one launch: cloud of the generated
 * sweep -> points inside the edited box (points_in_bbox_corners, one box [8, 3] per sample) -> pred_mask; paste where
 * pred_mask | gt_mask; range_pred [batch, 4, H*W] = (depth, intensity, pitch, yaw); edited cloud
 * points [batch, H*W, 5] = (x, y, z, intensity, beam index) in pixel order with count [batch]. */
typedef struct {
    const float* sample_depth;
    const float* sample_int;
    const float* depth_orig;
    const float* int_orig;
    const float* gt_mask; /* f32, non-zero = object (range_instance_mask_orig) */
    const float* bbox;    /* f32 [batch, 8, 3] */
    const float* pitch;
    const float* yaw;
    float* range_pred;
    uint8_t* pred_mask;
    float* points;
    int32_t* count;
    int32_t batch, H, W;
    float depth_min, depth_max;
} mobi_range_composite_args;
int mobi_range_composite(const mobi_range_composite_args* args, void* stream);

/* points_in_bbox_corners (ldm/data/box_np_ops.py:453-471, 406-427, 712-771): points f32 [n, point_stride >= 3],
 * corners f32 [m, 8, 3] in center_to_corner_box3d order -> out u8 [n, m] (1 = inside every face). */
int mobi_points_in_boxes(const float* points, int32_t point_stride, const float* corners, uint8_t* out, int32_t n, int32_t m,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOBI_B200_H */
