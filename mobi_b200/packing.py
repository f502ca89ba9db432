"""Weight packing: reference layouts (nn.Linear [out,in], nn.Conv2d OIHW) -> what the kernels read."""
import torch


def interleave_geglu(w, b):
    """GEGLU.proj (attention.py:41) produces [value | gate] halves; the GEGLU epilogue wants them interleaved
    in groups of 8 columns: [v0..7, g0..7, v8..15, g8..15, ...]."""
    n2 = w.shape[0]
    n = n2 // 2
    assert n % 8 == 0
    idx = torch.arange(n2, device=w.device).reshape(2, n // 8, 8).permute(1, 0, 2).reshape(-1)
    return w[idx].contiguous(), (b[idx].contiguous() if b is not None else None)


def interleave_geglu_pairs(w, b):
    """[value | gate] halves -> (value, gate) column pairs [v0, g0, v1, g1, ...] for MOBI_EPI_GEGLU2."""
    n2 = w.shape[0]
    n = n2 // 2
    idx = torch.arange(n2, device=w.device).reshape(2, n).t().reshape(-1)
    return w[idx].contiguous(), (b[idx].contiguous() if b is not None else None)


def pack_conv_weight(w, kpad=None):
    """OIHW -> [O, kh*kw*I] with K ordered (kh, kw, c), zero padded to kpad, bf16."""
    o, i, kh, kw = w.shape
    wp = w.permute(0, 2, 3, 1).reshape(o, kh * kw * i)
    if kpad is not None and kpad > wp.shape[1]:
        wp = torch.nn.functional.pad(wp, (0, kpad - wp.shape[1]))
    return wp.to(torch.bfloat16).contiguous()
