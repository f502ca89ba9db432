"""TEST INFRASTRUCTURE ONLY (oracle) — the bf16-operand floor of one UNet evaluation.

Runs the fp32 oracle (oracle/unet_oracle.py) with the operands of selected contraction classes (Linear, conv, QK^T, PV)
rounded to bf16 / fp16 and everything else (accumulation, norms, softmax, residuals) left in fp32: the error of that run
against the plain fp32 oracle is what ANY implementation that feeds bf16 operands to fp32-accumulating tensor cores
carries, however exact the rest of it is.  Parity tests quote it beside the measured error of the CUDA path, and
tools/precision_probe.py prints the per-class breakdown.
"""
import torch
import torch.nn.functional as RealF

from . import unet_oracle as uo


class _RoundingFunctional:
    """torch.nn.functional proxy that rounds GEMM / conv operands."""

    def __init__(self, lin=None, conv=None):
        self.lin, self.conv = lin, conv

    def __getattr__(self, n):
        return getattr(RealF, n)

    def linear(self, x, w, b=None):
        if self.lin is not None:
            x, w = x.to(self.lin).float(), w.to(self.lin).float()
        return RealF.linear(x, w, b)

    def conv2d(self, x, w, b=None, **kw):
        if self.conv is not None:
            x, w = x.to(self.conv).float(), w.to(self.conv).float()
        return RealF.conv2d(x, w, b, **kw)


class _RoundingTorch:
    def __init__(self, qk=None, pv=None):
        self.qk, self.pv = qk, pv

    def __getattr__(self, n):
        return getattr(torch, n)

    def einsum(self, eq, a, b):
        d = self.qk if eq == "bid,bjd->bij" else self.pv
        if d is not None:
            a, b = a.to(d).float(), b.to(d).float()
        return torch.einsum(eq, a, b)


def unet_forward_rounded(sd, cfg, x, t, c, lin=None, conv=None, qk=None, pv=None):
    """uo.unet_forward with the named operand classes rounded to the given dtypes (None = fp32)."""
    uo.F, uo.torch = _RoundingFunctional(lin, conv), _RoundingTorch(qk, pv)
    try:
        with torch.no_grad():
            return uo.unet_forward(sd, cfg, x, t, c)
    finally:
        uo.F, uo.torch = RealF, torch


def bf16_operand_floor(sd, cfg, x, t, c, ref=None, chunk=8):
    """max-abs-rel and relative-RMS error of the all-bf16-operand oracle against the fp32 oracle on (x, t, c)."""
    bf = torch.bfloat16
    outs, refs = [], []
    for i in range(0, x.shape[0], chunk):
        s = slice(i, i + chunk)
        outs.append(unet_forward_rounded(sd, cfg, x[s], t[s], c[s], lin=bf, conv=bf, qk=bf, pv=bf))
        if ref is None:
            with torch.no_grad():
                refs.append(uo.unet_forward(sd, cfg, x[s], t[s], c[s]))
    out = torch.cat(outs).double()
    ref = (torch.cat(refs) if ref is None else ref).double()
    return ((out - ref).abs().max() / ref.abs().max()).item(), ((out - ref).norm() / ref.norm()).item()


class _RoundingTorchBmm:
    def __init__(self, dt):
        self.dt = dt

    def __getattr__(self, n):
        return getattr(torch, n)

    def bmm(self, a, b):
        return torch.bmm(a.to(self.dt).float(), b.to(self.dt).float())


def vae_decode_rounded(sd, cfg, z, dt=torch.bfloat16):
    """oracle/vae_oracle.py:vae_decode with every conv / attention-product operand rounded to `dt` (fp32 accumulate)."""
    from . import vae_oracle as vo
    vo.F, vo.torch = _RoundingFunctional(conv=dt), _RoundingTorchBmm(dt)
    try:
        with torch.no_grad():
            return vo.vae_decode(sd, cfg, z)
    finally:
        vo.F, vo.torch = RealF, torch


def vae_bf16_operand_floor(sd, cfg, z, ref):
    out = vae_decode_rounded(sd, cfg, z).double()
    ref = ref.double()
    return ((out - ref).abs().max() / ref.abs().max()).item(), ((out - ref).norm() / ref.norm()).item()
