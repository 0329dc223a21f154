"""fp64 restatements of the one-token projection stages with the reference's bf16 rounding points
(model.py:376-387 wqkv + RoPE, 431/317 wo + residual, 442-443 FFN, 446-457 RMSNorm, 290-291 head).
Inputs are bf16 tensors (any device); results are bf16 tensors computed from exact (fp64) dot products, so a
kernel that accumulates in fp32 in any order must land on the same or the adjacent bf16 value."""
from __future__ import annotations

import torch


def bf(x):
    return x.to(torch.bfloat16)


def rmsnorm(x, w, eps):
    """bf16(bf16(x * rsqrt(mean(x^2) + eps)) * w), statistics in fp32 as RMSNorm.forward does (model.py:452-457)."""
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return bf(bf(xf * rstd).float() * w.float())


def matvec(w, x):
    return w.double() @ x.double()


def qkv_rope(x, norm_w, eps, wqkv, bias, freqs_row, n_q, n_kv, D):
    """x bf16 [dim] -> (q|k|v) bf16 [(n_q + 2 n_kv) * D], RoPE on q and k in fp32 (model.py:507-519)."""
    xn = rmsnorm(x, norm_w, eps)
    out = matvec(wqkv, xn)
    if bias is not None:
        out = out + bias.double()
    out = bf(out.float())   # nn.Linear output is a bf16 tensor (fp64 -> fp32 -> bf16: double rounding is below 2^-24 relative)
    rot = out[: (n_q + n_kv) * D].float().view(-1, D // 2, 2)
    c, s = freqs_row[:, 0].float(), freqs_row[:, 1].float()
    re = rot[..., 0] * c - rot[..., 1] * s
    im = rot[..., 1] * c + rot[..., 0] * s
    return torch.cat([bf(torch.stack([re, im], -1).reshape(-1)), out[(n_q + n_kv) * D:]])


def proj(w, x):
    return bf(matvec(w, x).float())


def ffn_gate(h, norm_w, eps, w1, w3):
    hn = rmsnorm(h, norm_w, eps)
    g = bf(matvec(w1, hn).float())
    u = bf(matvec(w3, hn).float())
    return bf(bf(torch.nn.functional.silu(g.float())).float() * u.float())


def ulp_excess(got, want):
    """max over elements of |got - want| / max(1 bf16 ulp at that magnitude, 1e-3 * rms(want)); <= 1.0 passes
    (an element that cancels to ~0 carries the absolute rounding error of the whole sum)."""
    a, b = got.float().cpu(), want.float().cpu()
    e = torch.floor(torch.log2(torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -126)))
    ulp = torch.pow(2.0, e - 7)
    allow = torch.maximum(ulp, 1e-3 * b.pow(2).mean().sqrt())
    return float(((a - b).abs() / allow).max())
