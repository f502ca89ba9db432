"""Training step of MObI's latent-diffusion UNet on the C-ABI CUDA kernels (BASELINE.json config 5).

What the reference runs per step (ldm/models/diffusion/ddpm.py):
  LatentDiffusion.forward (1040-1058): t ~ randint, conditioning (dropout of the whole batch with p = u_cond_percent),
  p_losses (1177-1217): noise ~ randn, x_noisy = cat(q_sample(x[:, :4], t, noise), x[:, 4:]), eps = apply_model(...),
  loss = mean((eps - noise)^2)  (logvar = 0, l_simple_weight = 1, original_elbo_weight = 0);
  loss.backward() through UNetModel with requires_grad only on parameters whose name contains "cond_adapter", "lidar"
  or "cross_modal" (DiffusionWrapper.__init__, 1686-1698); DDP all-reduce of those gradients; AdamW (1655).

Here the same computation is spelled out by hand, module by module: a training-mode forward that keeps what the backward
needs (the inference path's algebraic folds of the adapters are NOT used, because their weights are the trainable
ones), and a backward that mirrors it.  Every contraction is a tcgen05 GEMM / implicit conv (`ops.gemm`,
`ops.conv_implicit`): dgrad with transposed / flipped weight packs, wgrad as dY^T X over the token dimension, the
attention backward as five products per (row, head) around one softmax-backward kernel.  There is no autograd, no
torch math on activations and no CPU fallback.

The trainable parameters live in ONE flat fp32 buffer (their nn.Parameters are views into it), with a matching flat
gradient buffer: the data-parallel exchange is a single NCCL all-reduce over that buffer and the optimizer a single
fused AdamW launch.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from . import train_ops as tops
from .attention import LOG2E, BasicTransformerBlock, SpatialTransformer
from .openaimodel import Conv3x3, Downsample, ResBlock, Upsample
from .packing import pack_conv_weight

TRAINABLE_KEYS = ("cond_adapter", "lidar", "cross_modal")  # ddpm.py:1686-1698
BBOX_PREFIX = "cond_stage_model.bbox_embedder."             # checkpoint prefix of the trainable conditioning MLP


def is_trainable(name):
    return any(k in name for k in TRAINABLE_KEYS)


# ------------------------------------------------------------------------------------------------ flat parameter store
class FlatParams:
    """Moves the selected parameters of `module` into one flat f32 buffer (views stay registered as the module's
    nn.Parameters, so state_dict()/load_state_dict() keep working) and allocates the matching gradient buffer."""

    ALIGN = 8  # elements: every f32 view starts on a 32-byte boundary, so the bf16 mirror views at the same element
    #            offsets (UNetTrainer._pack_matrix) start on 16 bytes: what TMA / tcgen05 operand bases need

    def __init__(self, module, select=is_trainable, device=None, extra=()):
        """extra: further (name, parameter) pairs appended after the selected parameters of `module` (the trainable
        bbox_embedder of the conditioning stage, ddpm.py:580-586)."""
        named = [(n, p) for n, p in module.named_parameters() if select(n)] + list(extra)
        if not named:
            raise RuntimeError("no trainable parameters selected")
        device = device or named[0][1].device
        self.names = [n for n, _ in named]
        self.offsets, self.shapes = {}, {}
        off = 0
        for n, p in named:
            self.offsets[n] = off
            self.shapes[n] = tuple(p.shape)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.params = torch.zeros(off, device=device, dtype=torch.float32)
        # the gradient buffer carries 8 trailing floats: per-segment "received a gradient this step" flags that ride the
        # gradient all-reduce (UNetTrainer.step: a segment is active when ANY rank gave it a gradient, as under DDP)
        self.grads_full = torch.zeros(off + 8, device=device, dtype=torch.float32)
        self.grads = self.grads_full[:off]
        self.flags = self.grads_full[off:]
        with torch.no_grad():
            for n, p in named:
                v = self.view(self.params, n)
                v.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = v
                p.requires_grad_(True)
        for n, p in module.named_parameters():
            if not select(n):
                p.requires_grad_(False)

    def view(self, flat, name):
        o = self.offsets[name]
        shape = self.shapes[name]
        return flat[o:o + math.prod(shape)].view(shape)

    def grad(self, name):
        return self.view(self.grads, name)

    def pair_view(self, flat, first, second):
        """One [2*rows, cols] view over two adjacent equal-shape matrices (to_k | to_v)."""
        s = self.shapes[first]
        assert self.shapes[second] == s and self.offsets[second] == self.offsets[first] + math.prod(s), \
            "to_k / to_v are not adjacent in the flat buffer"
        o = self.offsets[first]
        return flat[o:o + 2 * math.prod(s)].view(2 * s[0], s[1])


def allreduce_mean_(flat, group=None):
    """Data-parallel gradient exchange: ONE all-reduce over the flat gradient buffer (NCCL over NVLink on the GPU box,
    gloo in the CPU tests), then the mean.  Zero gradients of parameters unused this step (bbox_uncond_vector off the
    cond-dropout steps) simply stay zero."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / world)
    return flat


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _conv_dgrad_pack(conv):
    """dX of a stride-1 'same' conv is the conv of dY with the spatially flipped, in/out-transposed filter."""
    w = conv.weight.detach().float()
    o, i, kh, kw = w.shape
    wt = w.permute(1, 0, 2, 3).flip(2, 3).contiguous()  # [I, O, kh, kw]
    k = kh * kw * o
    return dict(w=pack_conv_weight(wt, (k + 7) // 8 * 8), b=None, kh=kh, kw=kw, cin=o, cout=i, stride=1,
                pad=(kh // 2, kw // 2))


class _MnMajorB:
    """Tag: a forward weight pack [n_out, k_in] to be read as the MN-major B operand of dx = dy W."""
    __slots__ = ("w",)

    def __init__(self, w):
        self.w = w


class UNetTrainer:
    """forward_backward(x_start, t, noise, context) -> loss; step() -> all-reduce + AdamW + repack."""

    def __init__(self, ldm, lr=8e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, group=None, use_cuda_graph=True,
                 lse_backward=False, bbox_embedder=None, flash_backward=True, overlap_allreduce=True):
        self.ldm = ldm
        self.unet = ldm.model.diffusion_model if hasattr(ldm, "model") else ldm
        unet = self.unet
        if not (unet.multimodal and all(b.bbox_cond for b in self._blocks())):
            raise NotImplementedError("UNetTrainer: the training step is built for the joint camera+lidar UNet with "
                                      "bbox_cond (configs/mobi_nusc_*.yaml), the only trained configuration")
        dev = next(unet.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("UNetTrainer runs on CUDA only (no CPU fallback)")
        self.device = dev
        # the reference also trains the bbox_embedder of the conditioning stage (ddpm.py:580-586, 1636-1641): its parameters
        # join the same flat buffer under their checkpoint names
        self.bbox_embedder = bbox_embedder
        extra = [] if bbox_embedder is None else [(BBOX_PREFIX + n, p) for n, p in bbox_embedder.named_parameters()]
        if bbox_embedder is not None and isinstance(getattr(ldm, "bbox_uncond_vector", None), nn.Parameter):
            extra.append(("bbox_uncond_vector", ldm.bbox_uncond_vector))      # ddpm.py:477, 1643-1645
        self.flat = FlatParams(unet, is_trainable, dev, extra=extra)
        self.exp_avg = torch.zeros_like(self.flat.params)
        self.exp_avg_sq = torch.zeros_like(self.flat.params)
        # torch.optim.AdamW skips parameters whose gradient is None and counts steps per parameter: the flat buffer is
        # [UNet adapters | bbox_embedder | bbox_uncond_vector]; the last two only receive gradients on some steps
        # (bbox given / conditioning dropout, ddpm.py:1052-1056), so each segment has its own device-side step counter
        b1 = min([self.flat.offsets[n] for n in self.flat.names if n.startswith(BBOX_PREFIX)], default=self.flat.numel)
        b2 = self.flat.offsets.get("bbox_uncond_vector", self.flat.numel)
        self._adam_bounds = (b1, b2)
        self._adam_steps = torch.zeros(4, device=dev, dtype=torch.int32)
        self._adam_state = torch.zeros(12, device=dev, dtype=torch.float32)
        self._flag_rows = torch.tensor([[1, 0, 0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0, 0], [1, 0, 1, 0, 0, 0, 0, 0],
                                        [1, 1, 1, 0, 0, 0, 0, 0]], device=dev, dtype=torch.float32)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = group
        self.steps = 0
        self.loss_sum = torch.zeros(1, device=dev, dtype=torch.float32)
        self._names = {id(p): n for n, p in unet.named_parameters()}
        self._names.update({id(p): n for n, p in extra})
        self._ws = {}
        self.use_cuda_graph = use_cuda_graph
        # lse_backward: take the softmax row statistics of the attention backward from the forward kernel's log-sum-exp and
        # Delta = rowsum(dO * O) (flash-attention style) instead of recomputing them from the S / dP tiles: 8 % faster
        # step, but the rows of dS then sum to ~1e-3 instead of ~1e-7, which the un-normalised partner tokens amplify in
        # the cross-modal to_k / to_v weight gradients (27 % max-abs error on one tensor at full width, whole-gradient
        # cosine unchanged at 0.99992).  Off by default: the parity bar is per tensor.
        self.lse_backward = lse_backward
        # flash_backward: dQ' / dK / dV of every attention with head_dim <= 128 and tokens % 128 == 0 come from the two
        # flash-style kernels (attn_bwd_flash.cu) after the exact statistics pass; off = score tiles (dS, P) through HBM +
        # three batched GEMMs (kept as the A/B reference and for other shapes)
        self.flash_backward = flash_backward
        self._fb_graphs, self._seg_start = {}, None
        # overlap_allreduce: under torch.distributed the gradient buckets are all-reduced on a side stream while the next
        # phase of the backward runs (what DDP's bucketed reducer does for the reference, main.py:509-510)
        self.overlap_allreduce = overlap_allreduce
        self._comm_stream, self._pending = None, []
        unet.invalidate()
        unet.pack()
        self._pack_frozen_backward()
        self.repack_trainable()

    # ------------------------------------------------------------------ module lists / names
    def _blocks(self):
        return [m for m in self.unet.modules() if isinstance(m, BasicTransformerBlock)]

    def _g(self, param):
        return self.flat.grad(self._names[id(param)])

    # ------------------------------------------------------------------ packs
    @torch.no_grad()
    def _pack_frozen_backward(self):
        """Transposed (Linear) and flipped (conv) packs of the FROZEN weights for dgrad: built once."""
        bp = {}
        u = self.unet
        for m in u.modules():
            if isinstance(m, ResBlock):
                d = dict(conv1=_conv_dgrad_pack(m.in_layers[2]), conv2=_conv_dgrad_pack(m.out_layers[3]))
                if not isinstance(m.skip_connection, nn.Identity):
                    d["w_skip_T"] = _bf16(m.skip_connection.weight.reshape(m.out_channels, m.channels).t())
                bp[id(m)] = d
            elif isinstance(m, SpatialTransformer):
                bp[id(m)] = dict(w_in_T=_bf16(m.proj_in.weight.reshape(m.inner_dim, m.in_channels).t()),
                                 w_out_T=_bf16(m.proj_out.weight.reshape(m.in_channels, m.inner_dim).t()))
            elif isinstance(m, BasicTransformerBlock):
                p = m._p
                bp[id(m)] = dict(w_qkv_T=p["w_qkv"].t().contiguous(), w_o_T=p["w_o"].t().contiguous(),
                                 w_ff1_T=p["w_ff1"].t().contiguous(), w_ff2_T=p["w_ff2"].t().contiguous(),
                                 w_o2_T=p["w_o2"].t().contiguous(), w_v2_T=p["w_v2"].t().contiguous())
            elif isinstance(m, Upsample):
                bp[id(m)] = dict(conv=_conv_dgrad_pack(m.conv))
            elif isinstance(m, Downsample):
                bp[id(m)] = dict(conv=_conv_dgrad_pack(m.op))
        bp["conv_out"] = _conv_dgrad_pack(u.out[2])
        self.bp = bp

    def _pack_matrix(self, w, scale=1.0, transposed=True):
        """f32 master weight [N, K] (a view of the flat parameter buffer) -> its bf16 operand pack: a VIEW of the flat bf16
        mirror at the same offset (refreshed for all weights by one segmented-cast launch, `scale` folded in), and the same
        view tagged for the dgrad GEMM dx = dy W, which reads it as an MN-major B operand (no transposed pack)."""
        off = (w.data_ptr() - self.flat.params.data_ptr()) // 4
        assert 0 <= off and off + w.numel() <= self.flat.numel and w.is_contiguous(), "not a view of the flat parameter buffer"
        assert off % 8 == 0, "bf16 operand views must start on a 16-byte boundary (FlatParams.ALIGN)"
        self._segments.append((off, w.numel(), float(scale)))
        fwd = self._mirror[off:off + w.numel()].view(w.shape)
        return fwd, (_MnMajorB(fwd) if transposed else None)

    @staticmethod
    def _dgemm(dy, w_t, **kw):
        """dx = dy W for a weight given either as a transposed K-major pack [k_in, n_out] (frozen weights, packed once)
        or as the forward pack [n_out, k_in] read MN-major (trainable weights)."""
        if isinstance(w_t, _MnMajorB):
            w = w_t.w
            return ops.gemm(dy, w, M=dy.numel() // dy.shape[-1], N=w.shape[1], K=w.shape[0], ldb=w.stride(0), b_mn=True, **kw)
        return ops.gemm(dy, w_t, **kw)

    @torch.no_grad()
    def repack_trainable(self):
        """bf16 operand packs of the TRAINABLE weights after every optimizer step: ONE launch casts the whole flat f32
        parameter buffer into its bf16 mirror (attention scales folded into the to_q segments); the packs are views of
        the mirror, so their addresses never change and the captured forward/backward graphs keep reading them."""
        if self._seg_start is None:
            self._build_packs()
        ops.Stats.launches += 1
        L.check(L.load().mobi_cast_bf16_segments(self.flat.params.data_ptr(), self._mirror.data_ptr(), self.flat.numel,
                                                   self._seg_start.data_ptr(), self._seg_scale.data_ptr(),
                                                   self._seg_start.numel(), L.stream()), "cast_bf16_segments")
        self.unet.mark_trainable_stale()

    def _build_packs(self):
        """Once: the pack views (self.tp) and the (start, scale) table of the segmented cast."""
        self._mirror = torch.empty(self.flat.numel, device=self.device, dtype=torch.bfloat16)
        self._segments = []
        self._repack_trainable()
        bounds = {0: 1.0}
        for off, n, sc in sorted(self._segments):
            bounds[off] = sc
            bounds.setdefault(off + n, 1.0)          # what follows a scaled matrix is cast unscaled unless it is a pack itself
        starts = sorted(b for b in bounds if b < self.flat.numel)
        assert all(b % 4 == 0 for b in starts)
        self._seg_start = torch.tensor(starts, device=self.device, dtype=torch.int64)
        self._seg_scale = torch.tensor([bounds[b] for b in starts], device=self.device, dtype=torch.float32)
        # the same table for the gradients: the flat gradient buffer shares the parameter offsets, and only the to_q
        # segments carry a scale (every other pack is cast with scale 1)
        self._gseg_scale = self._seg_scale.clone()
        # gradient buckets in the order the backward completes them (flat offsets follow module order: input_blocks,
        # middle_block, output_blocks, then the extras): A = middle + output blocks, B = input blocks of the two coarser
        # levels, C = the level-0 input blocks, D = bbox_embedder / bbox_uncond_vector + the activity flags
        names = [n for n in self.flat.names if not (n.startswith(BBOX_PREFIX) or n == "bbox_uncond_vector")]
        first = lambda pred: min([self.flat.offsets[n] for n in names if pred(n)], default=None)           # noqa: E731
        unet_end = self._adam_bounds[0]
        a_lo = first(lambda n: n.startswith("middle_block") or n.startswith("output_blocks"))
        b_lo = first(lambda n: n.startswith("input_blocks.") and int(n.split(".")[1]) >= 4)
        a_lo = unet_end if a_lo is None else a_lo
        b_lo = a_lo if b_lo is None else b_lo
        self._buckets = [(a_lo, unet_end), (b_lo, a_lo), (0, b_lo), (unet_end, self.flat.numel + 8)]
        self._bucket_scale = []
        for lo, hi in self._buckets[:3]:   # the q-gradient rescale of scale_segments, cut at the bucket bounds
            st = [lo] + [b for b in starts if lo < b < hi]
            sc = [bounds[max(b2 for b2 in starts if b2 <= b)] for b in st]
            self._bucket_scale.append((torch.tensor([b - lo for b in st], device=self.device, dtype=torch.int64),
                                       torch.tensor(sc, device=self.device, dtype=torch.float32)))

    def _repack_trainable(self):
        tp = {}
        for blk in self._blocks():
            sc = blk.d_head ** -0.5
            d = {}
            ca = blk.cond_adapter_attn
            d["a_wq"], d["a_wq_T"] = self._pack_matrix(ca.to_q.weight.data, sc)
            d["a_wk"], d["a_wk_T"] = self._pack_matrix(ca.to_k.weight.data)
            d["a_wv"], d["a_wv_T"] = self._pack_matrix(ca.to_v.weight.data)
            d["a_wo"], d["a_wo_T"] = self._pack_matrix(ca.to_out[0].weight.data)
            d["a_bo"] = ca.to_out[0].bias.data
            d["a_wc"], d["a_wc_T"] = self._pack_matrix(blk.cond_adapter_connector.weight.data)
            d["a_bc"] = blk.cond_adapter_connector.bias.data
            for m in ("camera", "lidar"):
                at = getattr(blk, "cross_modal_attn_" + m)
                cn = getattr(blk, "cross_modal_connector_" + m)
                d[m + "_wq"], d[m + "_wq_T"] = self._pack_matrix(at.to_q.weight.data, sc * LOG2E)
                kv = self.flat.pair_view(self.flat.params, self._names[id(at.to_k.weight)], self._names[id(at.to_v.weight)])
                d[m + "_wkv"], d[m + "_wkv_T"] = self._pack_matrix(kv)
                d[m + "_wo"], d[m + "_wo_T"] = self._pack_matrix(at.to_out[0].weight.data)
                d[m + "_bo"] = at.to_out[0].bias.data
                d[m + "_wc"], d[m + "_wc_T"] = self._pack_matrix(cn.weight.data)
                d[m + "_bc"] = cn.bias.data
            tp[id(blk)] = d
        if self.bbox_embedder is not None:
            be = self.bbox_embedder
            lin = [be.bbox_proj, be.second_linear[0], be.second_linear[2], be.second_linear[4]]
            tp["bbox"] = [self._pack_matrix(m.weight.data) + (m.bias.data,) for m in lin]
        self.tp = tp
        # the inference packs fold these weights: they are stale now
        self.unet.mark_trainable_stale()

    # ------------------------------------------------------------------ attention helpers
    def _attn_fwd(self, q, k, v, B, H, D, T):
        """Returns (o [B, T, C] bf16, lse [B*H, T] f32 or None): the row-major-V kernel also emits the log-sum-exp of
        every score row, which spares the backward its statistics pass over the T x T tiles."""
        if D <= 128:
            lse = torch.empty((B * H, T), device=q.device, dtype=torch.float32) if getattr(self, "lse_backward", True) else None
            return ops.attention(q, k, v, B, H, D, T, T, v_rowmajor=True, lse=lse), lse
        vt = tops.transpose(v, rows=T, cols=D, batch=B * H, in_batch_stride=T * D)
        return ops.attention(q, k, vt, B, H, D, T, T), None

    def _attn_ws(self, H, T, fused=False):
        ws = self._ws.get((H, T, fused))
        if ws is None:
            dev = self.device
            f32_tiles = (0,) if fused else (H, T, T)      # the fused score-tile kernel never materialises S / dP
            ws = dict(S=torch.empty(f32_tiles, device=dev, dtype=torch.float32),
                      dP=torch.empty(f32_tiles, device=dev, dtype=torch.float32),
                      dS=torch.empty((H, T, T), device=dev, dtype=torch.bfloat16),
                      dSt=torch.empty((0,) if fused else (H, T, T), device=dev, dtype=torch.bfloat16),
                      Pt=torch.empty((H, T, T), device=dev, dtype=torch.bfloat16),
                      stats=torch.empty((3 * H * T,), device=dev, dtype=torch.float32))
            self._ws[(H, T, fused)] = ws
        return ws

    def _attn_bwd(self, q, k, v, do, B, H, D, T, dq_out, dq_col, dkv_out, dk_col, dv_col, o=None, lse=None):
        """Backward of softmax(q' k^T) v per (row, head) (CrossAttention.forward, attention.py:179-192).
        q, k, v: bf16 [B*H, T, D] (q' carries scale*log2e); do: bf16 [B*T, C] token-major.  Writes dq', dk, dv as
        [T, D] blocks at the given column offsets of the token-major outputs.  Three routes:
          * head_dim <= 128 and T % 128 == 0 (64^2 and 32^2 levels): exact statistics pass + the two flash-style kernels,
            no T x T tile in HBM (attn_bwd_flash.cu);
          * small levels (16^2, 8^2; head_dim 160): materialised f32 S / dP and bf16 dS / dS^T / P^T tiles, every product
            ONE launch over all batch rows and heads through the two-level batch of mobi_gemm;
          * otherwise (flash_backward=False, lse_backward=True, huge tiles): the same per batch row with a reused workspace,
            the score tiles from the fused score-tile kernel when possible."""
        C = H * D
        TD, TT = T * D, T * T
        fused = D <= 128 and T % 128 == 0 and not (lse is not None and getattr(self, "lse_backward", False))
        if fused and getattr(self, "flash_backward", True):
            # statistics pass + two flash-style kernels: dQ' (CTA = 128 query rows) and dK / dV (CTA = 128 key rows)
            # accumulate in TMEM from dS / P tiles staged in shared memory; no T x T tile is written to HBM
            stats = self._ws.get(("stats", B * H, T))
            if stats is None:
                stats = self._ws[("stats", B * H, T)] = torch.empty((3 * B * H * T,), device=q.device, dtype=torch.float32)
            tops.attn_bwd_flash(q, k, v, do, stats, dq_out[:, dq_col:dq_col + C], dkv_out[:, dk_col:dk_col + C],
                                dkv_out[:, dv_col:dv_col + C], heads=H, tokens=T, head_dim=D, dscale=tops.LN2, batch_rows=B)
            return
        # fused: the score tiles of ALL batch rows come from one launch pair (more (head, query-tile) items per wave)
        ws_all = self._attn_ws(B * H, T, True) if fused else None
        ws = ws_all if fused else self._attn_ws(H, T, False)
        f32 = torch.float32
        if fused:
            # S and dP are recomputed on tcgen05 inside the statistics pass and the main pass: the f32 score tiles never
            # reach HBM, only dS and P (bf16, both row-major) are written
            tops.attn_bwd_tiles(q, k, v, do, ws_all["stats"], ws_all["dS"], None, ws_all["Pt"], heads=H, tokens=T,
                                head_dim=D, ld_do=C, dscale=tops.LN2, batch_rows=B)
        if not fused and lse is None and B * H * TT * 14 <= (256 << 20):
            # small levels (16 x 16, 8 x 8: head_dim 160): ALL batch rows and heads in one launch per product through the
            # two-level batch of mobi_gemm (head h of batch row b of a token-major matrix = b * T * ld + h * D)
            wsa = self._attn_ws(B * H, T, False)
            two = dict(batch=B * H, batch_inner=H)
            ldq, ldkv = dq_out.stride(0), dkv_out.stride(0)
            ops.gemm(q, k, out=wsa["S"], out_dtype=f32, M=T, N=T, K=D, lda=D, ldb=D, ldo=T, batch=B * H, a_batch_stride=TD,
                     b_batch_stride=TD, out_batch_stride=TT)
            ops.gemm(do, v, out=wsa["dP"], out_dtype=f32, M=T, N=T, K=D, lda=C, ldb=D, ldo=T, a_batch_stride=D,
                     a_batch2_stride=T * C, b_batch_stride=TD, b_batch2_stride=H * TD, out_batch_stride=TT,
                     out_batch2_stride=H * TT, **two)
            tops.attn_softmax_bwd(wsa["S"], wsa["dP"], wsa["dS"], wsa["dSt"], wsa["Pt"], wsa["stats"], T, T, tops.LN2,
                                  batch=B * H)
            for a_op, b_op, ldb_, bs1, bs2, out, ldo_ in (
                    (wsa["dS"], k, D, TD, H * TD, dq_out[:, dq_col:dq_col + C], ldq),
                    (wsa["dSt"], q, D, TD, H * TD, dkv_out[:, dk_col:dk_col + C], ldkv),
                    (wsa["Pt"], do, C, D, T * C, dkv_out[:, dv_col:dv_col + C], ldkv)):
                ops.gemm(a_op, b_op, out=out, M=T, N=D, K=T, lda=T, ldb=ldb_, ldo=ldo_, b_mn=True, a_batch_stride=TT,
                         a_batch2_stride=H * TT, b_batch_stride=bs1, b_batch2_stride=bs2, out_batch_stride=D,
                         out_batch2_stride=T * ldo_, **two)
            return
        for b in range(B):
            rows = slice(b * T, (b + 1) * T)
            hs = slice(b * H, (b + 1) * H)
            bat = dict(batch=H)
            if fused:
                ws = {kk: (vv[hs] if kk != "stats" else vv) for kk, vv in ws_all.items()}
            else:
                ops.gemm(q[hs], k[hs], out=ws["S"], out_dtype=f32, M=T, N=T, K=D, lda=D, ldb=D, ldo=T, a_batch_stride=TD,
                         b_batch_stride=TD, out_batch_stride=TT, **bat)
                ops.gemm(do[rows], v[hs], out=ws["dP"], out_dtype=f32, M=T, N=T, K=D, lda=C, ldb=D, ldo=T,
                         a_batch_stride=D, b_batch_stride=TD, out_batch_stride=TT, **bat)
                if lse is not None:
                    tops.attn_softmax_bwd_lse(ws["S"], ws["dP"], ws["dS"], ws["dSt"], ws["Pt"], ws["stats"], T, T, tops.LN2,
                                              lse[hs], o[rows], do[rows], C, D, batch=H)
                else:
                    tops.attn_softmax_bwd(ws["S"], ws["dP"], ws["dS"], ws["dSt"], ws["Pt"], ws["stats"], T, T, tops.LN2,
                                          batch=H)
            # dq' = dS k ; dk = dS^T q' ; dv = P^T dO.  k, q' and the head columns of dO are read as MN-major B operands
            # (no transposed copies); the fused path also reads dS / P as MN-major A operands instead of writing dS^T / P^T
            a_dk, a_dv = (ws["dS"], ws["Pt"]) if fused else (ws["dSt"], ws["Pt"])
            mn = dict(a_mn=True) if fused else {}
            ops.gemm(ws["dS"], k[hs], out=dq_out[rows, dq_col:dq_col + C], M=T, N=D, K=T, lda=T, ldb=D, b_mn=True,
                     a_batch_stride=TT, b_batch_stride=TD, out_batch_stride=D, **bat)
            ops.gemm(a_dk, q[hs], out=dkv_out[rows, dk_col:dk_col + C], M=T, N=D, K=T, lda=T, ldb=D, b_mn=True,
                     a_batch_stride=TT, b_batch_stride=TD, out_batch_stride=D, **bat, **mn)
            ops.gemm(a_dv, do[rows], out=dkv_out[rows, dv_col:dv_col + C], M=T, N=D, K=T, lda=T, ldb=C, b_mn=True,
                     a_batch_stride=TT, b_batch_stride=D, out_batch_stride=D, **bat, **mn)

    # ------------------------------------------------------------------ BasicTransformerBlock
    def _block_forward(self, blk, x0, R, T, ctx_bf, ctx_f32):
        """attention.py:230-266, x0: f32 [R*T, C] (kept).  Returns (x6, tape)."""
        p, tp = blk._p, self.tp[id(blk)]
        C, H, D = blk.dim, blk.n_heads, blk.d_head
        dev, bf, f32 = x0.device, torch.bfloat16, torch.float32
        nk = ctx_f32.shape[1]
        t = dict(x0=x0)
        # 1. self-attention
        n1 = ops.layernorm(x0, *p["norm1"])
        q = torch.empty((R * H, T, D), device=dev, dtype=bf)
        k, v = torch.empty_like(q), torch.empty_like(q)
        ops.gemm(n1, p["w_qkv"], epilogue=L.EPI_QKV_ROW, heads=H, head_dim=D, tokens=T, out=q, out2=k, out3=v)
        o1, lse1 = self._attn_fwd(q, k, v, R, H, D, T)
        o1 = o1.reshape(R * T, C)
        x2 = ops.gemm(o1, p["w_o"], bias=p["b_o"], residual=x0, out_dtype=f32)
        t.update(q=q, k=k, v=v, o=o1, lse=lse1)
        # 2. attn2: one key -> the same vector to_out(to_v(c0)) for every token of a batch row (frozen weights)
        c0 = ctx_bf.reshape(R, nk, -1)[:, 0].contiguous()
        vec2 = ops.gemm(ops.gemm(c0, p["w_v2"]), p["w_o2"], bias=p["b_o2"], out_dtype=f32)
        # 3. bbox / reference adapter (trainable): LN -> to_q ; to_k, to_v on the context ; 2-key attention ; to_out ;
        #    connector.  The LayerNorm kernel adds vec2 to x2 in place first.
        lna = blk.cond_adapter_norm
        na = ops.layernorm(x2, lna.weight.data, lna.bias.data, add_vec=vec2, add_rows_per_vec=T)
        qa = ops.gemm(na, tp["a_wq"])
        ka = ops.gemm(ctx_bf, tp["a_wk"], out_dtype=f32).reshape(R, nk, C)
        va = ops.gemm(ctx_bf, tp["a_wv"], out_dtype=f32).reshape(R, nk, C)
        oa = tops.ctx_attn_qspace(qa, ka, va, R, T, H)
        ua = ops.gemm(oa, tp["a_wo"], bias=tp["a_bo"])
        x3 = ops.gemm(ua, tp["a_wc"], bias=tp["a_bc"], residual=x2, out_dtype=f32)
        t.update(x2=x2, na=na, qa=qa, ka=ka, va=va, oa=oa, ua=ua, x3=x3)
        # 4. cross-modal attention: camera rows (even) attend to the lidar rows, then lidar rows (odd) to the UPDATED
        #    camera rows (attention.py:246-263)
        Rh = R // 2
        x5 = x3.clone()
        seg = dict(rows=Rh * T, seg=T, seg_stride=2 * T)
        for m, off in (("camera", 0), ("lidar", T)):
            lnm = getattr(blk, "cross_modal_norm_" + m)
            nq = ops.layernorm(x5, lnm.weight.data, lnm.bias.data, seg_offset=off, **seg)
            ctx = ops.layernorm(x5, None, None, seg_offset=T - off, **seg)
            qm = torch.empty((Rh * H, T, D), device=dev, dtype=bf)
            km, vm = torch.empty_like(qm), torch.empty_like(qm)
            ops.gemm(nq, tp[m + "_wq"], epilogue=L.EPI_HEADS, heads=H, head_dim=D, tokens=T, out=qm)
            ops.gemm(ctx, tp[m + "_wkv"], epilogue=L.EPI_KV_ROW, heads=H, head_dim=D, tokens=T, out=km, out2=vm)
            om, lsem = self._attn_fwd(qm, km, vm, Rh, H, D, T)
            om = om.reshape(Rh * T, C)
            um = ops.gemm(om, tp[m + "_wo"], bias=tp[m + "_bo"])
            ops.gemm(um, tp[m + "_wc"], bias=tp[m + "_bc"], residual=x5, out=x5, ldo=C, out_seg=T, out_seg_stride=2 * T,
                     out_seg_offset=off)
            t[m] = dict(nq=nq, ctx=ctx, q=qm, k=km, v=vm, o=om, u=um, lse=lsem)
        # 5. GEGLU feed-forward
        n3 = ops.layernorm(x5, *p["norm3"])
        g = ops.gemm(n3, p["w_ff1"], bias=p["b_ff1"])
        x6 = ops.gemm(tops.geglu(g), p["w_ff2"], bias=p["b_ff2"], residual=x5, out_dtype=f32)
        t.update(x5=x5, g=g)
        return x6, t

    def _linear_bwd(self, dy_bf, x_bf, w_T, weight, bias, M):
        """y = x W^T + b: accumulates dW, db and returns dx = dy W (bf16)."""
        n_out, k_in = weight.shape
        if bias is not None:
            tops.colsum(dy_bf, self._g(bias).reshape(1, n_out))
        tops.wgrad(dy_bf, x_bf, self._g(weight), M=M, n_out=n_out, k_in=k_in)
        return self._dgemm(dy_bf, w_T)

    def _block_backward(self, blk, t, d, R, T, ctx_f32, need_dx0=True):
        """d: f32 [R*T, C] gradient w.r.t. the block output; turned into the gradient w.r.t. x0 in place."""
        p, tp, bp = blk._p, self.tp[id(blk)], self.bp[id(blk)]
        C, H, D = blk.dim, blk.n_heads, blk.d_head
        dev, bf = d.device, torch.bfloat16
        M, Rh = R * T, R // 2
        Mh = Rh * T
        nk = ctx_f32.shape[1]
        # 5. feed-forward
        dh = ops.gemm(ops.cast_bf16(d), bp["w_ff2_T"])
        dn3 = ops.gemm(tops.geglu_bwd(t["g"], dh), bp["w_ff1_T"])
        tops.layernorm_bwd(t["x5"], p["norm3"][0], dn3, d)
        # 4. cross-modal, in reverse order: lidar first (its context gradient lands on the camera rows)
        for m, off in (("lidar", T), ("camera", 0)):
            s = t[m]
            at = getattr(blk, "cross_modal_attn_" + m)
            cn = getattr(blk, "cross_modal_connector_" + m)
            ln = getattr(blk, "cross_modal_norm_" + m)
            seg = dict(seg=T, seg_stride=2 * T)
            dm = ops.layernorm(d, None, None, rows=Mh, seg_offset=off, **seg)  # bf16 copy of this modality's rows
            du = self._linear_bwd(dm, s["u"], tp[m + "_wc_T"], cn.weight, cn.bias, Mh)
            do = self._linear_bwd(du, s["o"], tp[m + "_wo_T"], at.to_out[0].weight, at.to_out[0].bias, Mh)
            dq = torch.empty((Mh, C), device=dev, dtype=bf)
            dkv = torch.empty((Mh, 2 * C), device=dev, dtype=bf)
            self._attn_bwd(s["q"], s["k"], s["v"], do, Rh, H, D, T, dq, 0, dkv, 0, C, o=s["o"], lse=s["lse"])
            tops.wgrad(dq, s["nq"], self._g(at.to_q.weight), M=Mh, n_out=C, k_in=C)
            dnq = self._dgemm(dq, tp[m + "_wq_T"])
            gkv = self.flat.pair_view(self.flat.grads, self._names[id(at.to_k.weight)], self._names[id(at.to_v.weight)])
            tops.wgrad(dkv, s["ctx"], gkv, M=Mh, n_out=2 * C, k_in=C)
            dctx = self._dgemm(dkv, tp[m + "_wkv_T"])
            # the LayerNorm input of this modality's rows is x3 (lidar rows are untouched by the camera update)
            tops.layernorm_bwd(t["x3"], ln.weight.data, dnq, d, rows=Mh, seg_offset=off, dgamma=self._g(ln.weight),
                               dbeta=self._g(ln.bias), **seg)
            tops.scatter_add_rows(dctx, d, rows=Mh, Cc=C, seg_offset=T - off, **seg)
        # 3. bbox / reference adapter
        ca, cn, ln = blk.cond_adapter_attn, blk.cond_adapter_connector, blk.cond_adapter_norm
        du = self._linear_bwd(ops.cast_bf16(d), t["ua"], tp["a_wc_T"], cn.weight, cn.bias, M)
        do = self._linear_bwd(du, t["oa"], tp["a_wo_T"], ca.to_out[0].weight, ca.to_out[0].bias, M)
        dqa, dka, dva = tops.ctx_attn_qspace(t["qa"], t["ka"], t["va"], R, T, H, d_o=do)
        tops.wgrad(dqa, t["na"], self._g(ca.to_q.weight), M=M, n_out=C, k_in=C)
        dna = self._dgemm(dqa, tp["a_wq_T"])
        cflat = ctx_f32.reshape(R * nk, -1)
        tops.wgrad_small(dka.reshape(R * nk, C), cflat, self._g(ca.to_k.weight))
        tops.wgrad_small(dva.reshape(R * nk, C), cflat, self._g(ca.to_v.weight))
        tops.layernorm_bwd(t["x2"], ln.weight.data, dna, d, dgamma=self._g(ln.weight), dbeta=self._g(ln.bias))
        # gradient w.r.t. the conditioning tokens (what trains the bbox_embedder): both tokens through the adapter's
        # to_k / to_v, token 0 also through the frozen attn2 (vec2 = to_out(to_v(c0)) added to every token of a row)
        dctx = self.d_context.reshape(R * nk, -1)
        self._dgemm(ops.cast_bf16(dka.reshape(R * nk, C)), tp["a_wk_T"], residual=dctx, out=dctx)
        self._dgemm(ops.cast_bf16(dva.reshape(R * nk, C)), tp["a_wv_T"], residual=dctx, out=dctx)
        dvec2 = torch.zeros((R, C), device=dev, dtype=torch.float32)
        tops.colsum(d, dvec2, rows_per_group=T)
        dv0 = ops.gemm(ops.cast_bf16(dvec2), bp["w_o2_T"])
        dc0 = self.d_context[:, 0]
        ops.gemm(dv0, bp["w_v2_T"], residual=dc0, out=dc0)
        # 2. attn2 adds a per-row constant: identity for d.   1. self-attention (frozen weights: dgrad only)
        if need_dx0:
            do1 = ops.gemm(ops.cast_bf16(d), bp["w_o_T"])
            dqkv = torch.empty((M, 3 * C), device=dev, dtype=bf)
            self._attn_bwd(t["q"], t["k"], t["v"], do1, R, H, D, T, dqkv, 0, dqkv, C, 2 * C, o=t["o"], lse=t["lse"])
            dn1 = ops.gemm(dqkv, bp["w_qkv_T"])
            tops.layernorm_bwd(t["x0"], p["norm1"][0], dn1, d)
        return d

    # ------------------------------------------------------------------ SpatialTransformer / ResBlock / resampling
    def _st_forward(self, st, h, ctx_bf, ctx_f32):
        p = st._p
        R, Hh, Ww, C = h.shape
        T = Hh * Ww
        hn = ops.groupnorm(h, p["gn"][0], p["gn"][1], 1e-6, silu=False)
        x = ops.gemm(hn.reshape(R * T, C), p["w_in"], bias=p["b_in"], out_dtype=torch.float32)
        tapes = []
        for blk in st.transformer_blocks:
            x, tb = self._block_forward(blk, x, R, T, ctx_bf, ctx_f32)
            tapes.append(tb)
        out = ops.gemm(ops.cast_bf16(x), p["w_out"], bias=p["b_out"], residual=h.reshape(R * T, C),
                       out_dtype=torch.float32)
        return out.reshape(R, Hh, Ww, C), dict(h=h, blocks=tapes)

    def _st_backward(self, st, t, dout, ctx_f32, need_dx=True):
        p, bp = st._p, self.bp[id(st)]
        R, Hh, Ww, C = dout.shape
        T = Hh * Ww
        d = ops.gemm(ops.cast_bf16(dout).reshape(R * T, C), bp["w_out_T"], out_dtype=torch.float32)
        n = len(st.transformer_blocks)
        for i in range(n - 1, -1, -1):
            d = self._block_backward(st.transformer_blocks[i], t["blocks"][i], d, R, T, ctx_f32,
                                     need_dx0=need_dx or i > 0)
        if not need_dx:
            return None
        dhn = ops.gemm(ops.cast_bf16(d), bp["w_in_T"])
        dh, _ = tops.groupnorm_bwd(t["h"], p["gn"][0], p["gn"][1], dhn, 1e-6, silu=False, dres=dout)
        return dh

    def _res_forward(self, rb, h, emb_out, skip):
        p = rb._p
        R, Hh, Ww, _ = h.shape
        need_raw = "w_skip" in p
        g = ops.groupnorm(h, p["gn1"][0], p["gn1"][1], 1e-5, x2=skip, silu=True, want_concat=need_raw)
        hn, raw = g if need_raw else (g, None)
        h1 = Conv3x3.run(p["conv1"], hn, row_bias=emb_out)
        hn2 = ops.groupnorm(h1, p["gn2"][0], p["gn2"][1], 1e-5, silu=True)
        if need_raw:
            res = ops.gemm(raw.reshape(R * Hh * Ww, -1), p["w_skip"], bias=p["b_skip"], out_dtype=torch.float32)
            res = res.reshape(R, Hh, Ww, rb.out_channels)
        else:
            res = h
        return Conv3x3.run(p["conv2"], hn2, residual=res), dict(h=h, skip=skip, h1=h1)

    def _res_backward(self, rb, t, dout):
        p, bp = rb._p, self.bp[id(rb)]
        R, Hh, Ww, Co = dout.shape
        dout_bf = ops.cast_bf16(dout)
        dc = Conv3x3.run(bp["conv2"], dout_bf)
        dh1, _ = tops.groupnorm_bwd(t["h1"], p["gn2"][0], p["gn2"][1], dc, 1e-5, silu=True)
        da = Conv3x3.run(bp["conv1"], ops.cast_bf16(dh1))
        if "w_skip_T" in bp:
            dres = ops.gemm(dout_bf.reshape(R * Hh * Ww, Co), bp["w_skip_T"], out_dtype=torch.float32)
        else:
            dres = dout
        return tops.groupnorm_bwd(t["h"], p["gn1"][0], p["gn1"][1], da, 1e-5, x2=t["skip"], silu=True, dres=dres)

    def _seq_forward(self, seq, h, skip, emb_all, ctx_bf, ctx_f32):
        tapes = []
        for layer in seq:
            if isinstance(layer, ResBlock):
                off = self.unet._p["emb_offs"][id(layer)]
                h, t = self._res_forward(layer, h, emb_all[:, off:off + layer.out_channels], skip)
                skip = None
            elif isinstance(layer, SpatialTransformer):
                h, t = self._st_forward(layer, h, ctx_bf, ctx_f32)
            elif isinstance(layer, Upsample):
                h, t = layer.run(h), None
            elif isinstance(layer, Downsample):
                h, t = layer.run(h), None
            else:
                raise RuntimeError("unexpected layer %s" % type(layer))
            tapes.append(t)
        return h, tapes

    def _seq_backward(self, seq, tapes, dh, ctx_f32, first=False):
        """Returns (dh, dskip).  `first`: input_blocks[1] — nothing trainable upstream of its transformer."""
        dskip = None
        layers = list(seq)
        for i in range(len(layers) - 1, -1, -1):
            layer, t = layers[i], tapes[i]
            if isinstance(layer, ResBlock):
                if first:
                    return None, None
                dh, dskip = self._res_backward(layer, t, dh)
            elif isinstance(layer, SpatialTransformer):
                dh = self._st_backward(layer, t, dh, ctx_f32, need_dx=not first)
            elif isinstance(layer, Upsample):
                dup = Conv3x3.run(self.bp[id(layer)]["conv"], ops.cast_bf16(dh))
                dh = tops.sum2x2(dup)
            elif isinstance(layer, Downsample):
                dh = Conv3x3.run(self.bp[id(layer)]["conv"], tops.zero_insert2x(dh))
        return dh, dskip

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def forward_backward(self, x_start, t, noise, context=None, *, bbox=None, uncond=False):
        """p_losses (ddpm.py:1177-1217) + backward.  x_start [R, 9, h, w] f32 (4 latent + 4 inpaint_image + mask channels),
        t int64 [R], noise [R, 4, h, w], context [R, n_ctx, ctx_dim].
        With a trainable bbox_embedder (constructor) and `bbox` [R, 8, 3] given, token 1 of the context is recomputed
        here from the box corners (get_learned_conditioning, ddpm.py:610-630) and its MLP receives gradients.
        uncond=True is the reference's whole-batch conditioning dropout (ddpm.py:1052-1056): the context becomes
        [learnable_vector, bbox_uncond_vector] for every row and bbox_uncond_vector receives its gradient.
        Gradients of the trainable parameters are left in self.flat.grads, the gradient w.r.t. the context in
        self.d_context; returns the loss (0-dim tensor).
        With use_cuda_graph the whole forward + backward (several thousand launches) is captured once per (input shapes,
        mode) and replayed from static input buffers."""
        if uncond:
            context, bbox = self._uncond_context(x_start.shape[0]), None
        for a in (x_start, noise, context, t) + ((bbox,) if bbox is not None else ()):
            if not a.is_cuda:
                raise RuntimeError("UNetTrainer needs CUDA tensors (no CPU fallback)")
        if bbox is not None and self.bbox_embedder is None:
            raise RuntimeError("forward_backward(bbox=...) needs UNetTrainer(bbox_embedder=...)")
        if not self.use_cuda_graph:
            return self._forward_backward(x_start, t, noise, context, bbox, uncond)
        key = (tuple(x_start.shape), tuple(noise.shape), tuple(context.shape), None if bbox is None else tuple(bbox.shape),
               bool(uncond))
        entry = self._fb_graphs.get(key)
        if entry is None:
            st = dict(x=x_start.detach().float().contiguous().clone(), t=t.to(torch.int64).contiguous().clone(),
                      noise=noise.detach().float().contiguous().clone(), ctx=context.detach().float().contiguous().clone(),
                      bbox=None if bbox is None else bbox.detach().float().contiguous().clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                # warm-up: workspaces, cudaFuncSetAttribute, allocator pools
                self._forward_backward(st["x"], st["t"], st["noise"], st["ctx"], st["bbox"], uncond)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            # one CUDA graph per phase of the backward (shared memory pool): between two replays the gradients of the
            # finished bucket go to the all-reduce on the communication stream
            gen = self._fb_phases(st["x"], st["t"], st["noise"], st["ctx"], st["bbox"], uncond)
            graphs, pool, loss = [], None, None
            before = ops.Stats.launches
            while loss is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    try:
                        next(gen)
                    except StopIteration as done:
                        loss = done.value
                pool = g.pool()
                graphs.append(g)
            entry = dict(graphs=graphs, static=st, loss=loss, d_context=self.d_context, kernels=ops.Stats.launches - before)
            self._fb_graphs[key] = entry
        st = entry["static"]
        st["x"].copy_(x_start)
        st["t"].copy_(t)
        st["noise"].copy_(noise)
        st["ctx"].copy_(context)
        if bbox is not None:
            st["bbox"].copy_(bbox)
        overlap = self._overlap_world() > 1
        self._pending = []
        for i, g in enumerate(entry["graphs"]):
            g.replay()
            if overlap:
                self._start_allreduce([i] if i < len(entry["graphs"]) - 1 else [i, 3])
        ops.Stats.launches += entry["kernels"]
        self.d_context = entry["d_context"]
        return entry["loss"]

    # ------------------------------------------------------------------ gradient exchange overlapped with the backward
    def _overlap_world(self):
        import torch.distributed as dist
        if not self.overlap_allreduce or not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def _start_allreduce(self, buckets):
        """SUM all-reduce of the given gradient buckets on the communication stream, ordered after everything the compute
        stream has issued so far; step() waits for them and folds the 1 / world of the mean into AdamW's gradient scale."""
        import torch.distributed as dist
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            for b in buckets:
                lo, hi = self._buckets[b]
                if hi > lo:
                    self._pending.append(dist.all_reduce(self.flat.grads_full[lo:hi], op=dist.ReduceOp.SUM,
                                                         group=self.group, async_op=True))

    def _uncond_context(self, R):
        """[learnable_vector, bbox_uncond_vector] repeated for every row (ddpm.py:1053-1056)."""
        ldm = self.ldm
        ctx = torch.empty((R, 2, ldm.learnable_vector.shape[-1]), device=self.device, dtype=torch.float32)
        ctx[:, 0].copy_(ldm.learnable_vector.detach().reshape(1, -1).expand(R, -1))
        ctx[:, 1].copy_(ldm.bbox_uncond_vector.detach().reshape(1, -1).expand(R, -1))
        return ctx

    # ------------------------------------------------------------------ trainable bbox_embedder (modules.py:181-213)
    def _bbox_forward(self, bbox):
        """bbox [R, 8, 3] -> (token f32 [R, 768], tape).  Same math as encoders.BBoxEmbedder.forward, keeping the
        pre-activations of the two SiLUs."""
        be, pk = self.bbox_embedder, self.tp["bbox"]
        R = bbox.shape[0]
        e = ops.fourier_embed(bbox.detach().float().contiguous(), be.num_freqs).reshape(R, 8 * be.out_dim)
        a0 = ops.gemm(e, pk[0][0], bias=pk[0][2])                       # bbox_proj (no activation)
        a1 = ops.gemm(a0, pk[1][0], bias=pk[1][2])                      # second_linear.0, SiLU follows
        h1 = ops.silu(a1)
        a2 = ops.gemm(h1, pk[2][0], bias=pk[2][2])                      # second_linear.2, SiLU follows
        h2 = ops.silu(a2)
        out = ops.gemm(h2, pk[3][0], bias=pk[3][2], out_dtype=torch.float32)
        return out, dict(e=e, a0=a0, a1=a1, h1=h1, a2=a2, h2=h2)

    def _bbox_backward(self, t, d_out):
        """d_out f32 [R, 768]: gradient w.r.t. the bbox token.  R is tiny (2 rows per joint sample): the weight gradients
        are plain outer-product sums (mobi_wgrad_small), the data gradients tensor-core GEMMs with 4-8 rows."""
        be, pk = self.bbox_embedder, self.tp["bbox"]
        lin = [be.bbox_proj, be.second_linear[0], be.second_linear[2], be.second_linear[4]]
        acts = [t["e"], t["a0"], t["h1"], t["h2"]]                       # inputs of the four Linears
        pre = [None, None, t["a1"], t["a2"]]                            # pre-activation feeding Linear i through a SiLU
        d = d_out.contiguous()
        for i in (3, 2, 1, 0):
            tops.wgrad_small(d, acts[i].float(), self._g(lin[i].weight))
            tops.colsum(d, self._g(lin[i].bias).reshape(1, -1))
            if i == 0:
                break
            dx = self._dgemm(ops.cast_bf16(d), pk[i][1], out_dtype=torch.float32)  # d @ W_i
            d = tops.silu_bwd(pre[i], dx) if pre[i] is not None else dx

    def _forward_backward(self, x_start, t, noise, context, bbox=None, uncond=False):
        """All phases back to back (eager mode / warm-up)."""
        gen = self._fb_phases(x_start, t, noise, context, bbox, uncond)
        while True:
            try:
                next(gen)
            except StopIteration as done:
                return done.value

    def _scale_bucket(self, i):
        """The q-gradient rescale (to_q packs carry the attention scale) of gradient bucket i, in place."""
        lo, hi = self._buckets[i]
        if hi <= lo:
            return
        st, sc = self._bucket_scale[i]
        ops.Stats.launches += 1
        L.check(L.load().mobi_scale_segments(self.flat.grads[lo:hi].data_ptr(), hi - lo, st.data_ptr(), sc.data_ptr(),
                                             st.numel(), L.stream()), "scale_segments")

    def _fb_phases(self, x_start, t, noise, context, bbox=None, uncond=False):
        """Generator over the three phases of one forward + backward; after phase i the gradients of bucket i
        (self._buckets) are final, so their all-reduce can run while the next phase computes:
          phase 0: forward, loss, backward of the output blocks and the middle block   -> bucket 0
          phase 1: backward of the input blocks of the coarser levels                  -> bucket 1
          phase 2: backward of the level-0 input blocks, bbox_embedder, uncond vector  -> buckets 2, 3
        Returns the loss (StopIteration.value)."""
        u, ldm = self.unet, self.ldm
        p = u._p
        R = x_start.shape[0]
        self.flat.grads.zero_()
        # which segments receive a gradient this step: the adapters always, the bbox_embedder when its token is computed
        # here, bbox_uncond_vector on conditioning-dropout steps (everything else is a None gradient for the optimizer)
        self.flat.flags.copy_(self._flag_rows[(1 if bbox is not None else 0) + (2 if uncond else 0)])
        self.loss_sum.zero_()
        t = t.to(torch.int64).contiguous()
        x_noisy = tops.q_sample(x_start.float().contiguous(), noise.float().contiguous(), ldm.sqrt_alphas_cumprod,
                                ldm.sqrt_one_minus_alphas_cumprod, t, noise.shape[1])
        ctx_f32 = context.detach().float().contiguous()
        bbox_tape = None
        if bbox is not None:   # token 1 = bbox_embedder(bbox), computed here so that its MLP can be trained
            tok, bbox_tape = self._bbox_forward(bbox)
            ctx_f32 = ctx_f32.clone()
            ctx_f32[:, 1].copy_(tok)
        self.d_context = torch.zeros_like(ctx_f32)
        ctx_bf = ops.cast_bf16(ctx_f32).reshape(R * ctx_f32.shape[1], -1)
        # ---- forward (openaimodel.py:861-898), keeping the tape
        t_emb = ops.timestep_embedding(t, u.model_channels)
        e1 = ops.gemm(t_emb, p["w_t0"], bias=p["b_t0"], act=1)
        e2 = ops.gemm(e1, p["w_t2"], bias=p["b_t2"], act=1)
        emb_all = ops.gemm(e2, p["w_emb"], bias=p["b_emb"], out_dtype=torch.float32)
        h = Conv3x3.run(p["conv_in"], ops.nchw_to_nhwc(x_noisy))
        hs, tapes_in, tapes_out = [h], [], []
        for seq in list(u.input_blocks)[1:]:
            h, tp_ = self._seq_forward(seq, h, None, emb_all, ctx_bf, ctx_f32)
            hs.append(h)
            tapes_in.append(tp_)
        h, tape_mid = self._seq_forward(u.middle_block, h, None, emb_all, ctx_bf, ctx_f32)
        for seq in u.output_blocks:
            h, tp_ = self._seq_forward(seq, h, hs.pop(), emb_all, ctx_bf, ctx_f32)
            tapes_out.append(tp_)
        hn = ops.groupnorm(h, p["gn_out"][0], p["gn_out"][1], 1e-5, silu=True)
        eps = Conv3x3.run(p["conv_out"], hn)                      # NHWC f32 [R, h, w, 4]
        # ---- loss and its gradient (mean over all elements)
        target = ops.nchw_to_nhwc(noise.float().contiguous())
        d_eps = tops.mse_grad(eps, target, self.loss_sum, 2.0 / eps.numel())
        self.eps_nhwc = eps
        # ---- backward
        dhn = Conv3x3.run(self.bp["conv_out"], ops.cast_bf16(d_eps))
        dh, _ = tops.groupnorm_bwd(h, p["gn_out"][0], p["gn_out"][1], dhn, 1e-5, silu=True)
        d_hs = []
        for seq, tp_ in zip(reversed(list(u.output_blocks)), reversed(tapes_out)):
            dh, dskip = self._seq_backward(seq, tp_, dh, ctx_f32)
            d_hs.append(dskip)                                     # gradients of hs[0], hs[1], ... in order
        dh, _ = self._seq_backward(u.middle_block, tape_mid, dh, ctx_f32)
        # gradients w.r.t. the scaled query projections -> w.r.t. to_q.weight, bucket by bucket
        self._scale_bucket(0)
        yield 0
        inputs = list(u.input_blocks)
        cut = False
        for i in range(len(inputs) - 1, 0, -1):
            if i == 3 and i != len(inputs) - 1:   # what is left are the level-0 blocks (input_blocks 1, 2): bucket 1 is final
                self._scale_bucket(1)
                cut = True
                yield 1
            dh = ops.add_f32(dh, d_hs[i])
            dh, _ = self._seq_backward(inputs[i], tapes_in[i - 1], dh, ctx_f32, first=(i == 1))
        if not cut:               # shallow (test) models with fewer than four input blocks
            self._scale_bucket(1)
            yield 1
        if bbox_tape is not None:
            self._bbox_backward(bbox_tape, self.d_context[:, 1])
        if uncond and "bbox_uncond_vector" in self.flat.offsets:   # every row saw the same vector: sum over rows
            tops.colsum(self.d_context[:, 1].contiguous(), self.flat.grad("bbox_uncond_vector").reshape(1, -1))
        self._scale_bucket(2)
        return self.loss_sum[0] / eps.numel()

    @torch.no_grad()
    def step(self, lr=None):
        """DDP gradient all-reduce (mean) + AdamW over the flat buffers + repack of the bf16 operand copies."""
        grad_scale = 1.0
        if self._pending:       # bucket all-reduces (SUM) started by forward_backward: wait, average inside AdamW
            for work in self._pending:
                work.wait()
            self._pending = []
            grad_scale = 1.0 / self._overlap_world()
        else:                   # gradients + the per-segment activity flags behind them
            allreduce_mean_(self.flat.grads_full, self.group)
        self.steps += 1
        tops.adamw_segments(self.flat.params, self.flat.grads, self.exp_avg, self.exp_avg_sq, bounds=self._adam_bounds,
                            flags=self.flat.flags, steps=self._adam_steps, state=self._adam_state,
                            lr=self.lr if lr is None else lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                            weight_decay=self.weight_decay, grad_scale=grad_scale)
        self.repack_trainable()           # also marks the UNet's inference packs of these weights stale
        if self.bbox_embedder is not None:
            self.bbox_embedder.invalidate()

    def named_grads(self):
        return {n: self.flat.grad(n) for n in self.flat.names}
