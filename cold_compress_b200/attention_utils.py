"""Attention wrappers around the cache (reference attention_utils.py:8-54).

* `scaled_dot_product_attention` keeps the reference signature for callers that hold explicit
  K/V tensors (prefill, or generic use).  It is plain tensor code: the decode step never comes
  through here -- `Attention.forward` hands decode to `KVCache.decode_step` (one CUDA kernel).
* `prefill_attention_stats` is the prefill path for caches that need attention probabilities:
  it produces y plus the two reductions of the pooled probability map that the prompt
  compressors and caches consume (`PrefillAttnStats`), block by block, so the [Hq, S, S] map the
  reference materialises (4 GiB at S=8192) never exists.  Rounding points follow the reference's
  naive path: bf16 after q@k^T, after *scale, after softmax, after the GQA mean.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch.nn import functional as F

from .prompt_compression import PrefillAttnStats


def scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0, scale=None,
                                 return_attn=False, attn_top_k=1.0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    n_q, n_kv = query.size(-2), key.size(-2)
    top_k = n_kv if n_q > 1 else int(attn_top_k * n_kv)
    if not return_attn and top_k == n_kv:
        return F.scaled_dot_product_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p,
                                              scale=scale), None
    s = 1.0 / math.sqrt(query.size(-1)) if scale is None else scale
    w = query @ key.transpose(-2, -1) * s
    if attn_mask is not None:
        assert top_k == n_kv, "Top-k attention not supported with masks."
        w = w + torch.zeros_like(w).masked_fill_(attn_mask.logical_not(), float("-inf"))
    if top_k < n_kv:
        _, keep = w.topk(top_k, dim=-1)
        value = value.gather(-2, keep.view(*keep.shape[:2], top_k, 1).expand(-1, -1, -1, value.shape[-1]))
        w = w.gather(-1, keep)
    p = torch.softmax(w, dim=-1)
    return p @ value, p


@torch.no_grad()
def prefill_attention_stats(q, k, v, obs_len: int = 16, block: int = 512, windows=()):
    """Causal self-attention over a prompt with GQA.

    q [1, Hq, S, D], k/v [1, Hkv, S, D] (bf16, post-RoPE).  Returns (y [1, Hq, S, D], PrefillAttnStats).
    Work per query block b: logits [Hkv, G*B, S_b] with S_b = end of the block (causality)."""
    _, Hq, S, D = q.shape
    Hkv = k.shape[1]
    G = Hq // Hkv
    scale = 1.0 / math.sqrt(D)
    obs_len = min(obs_len, S)
    qg = q[0].view(Hkv, G, S, D)
    kt = k[0].transpose(-2, -1)  # [Hkv, D, S]
    vv = v[0]
    y = torch.empty(Hkv, G, S, D, dtype=q.dtype, device=q.device)
    col_sum = torch.zeros(Hkv, S, dtype=torch.float32, device=q.device)
    obs_rows = torch.zeros(Hkv, obs_len, S, dtype=q.dtype, device=q.device)
    band = {int(w): torch.zeros(Hkv, S, dtype=torch.float32, device=q.device) for w in windows}
    cols = torch.arange(S, device=q.device)
    for i0 in range(0, S, block):
        i1 = min(S, i0 + block)
        B = i1 - i0
        logits = (qg[:, :, i0:i1].reshape(Hkv, G * B, D) @ kt[:, :, :i1]) * scale      # bf16, bf16
        rows = torch.arange(i0, i1, device=q.device).repeat(G).view(1, G * B, 1)
        logits = logits.masked_fill(cols[:i1].view(1, 1, -1) > rows, float("-inf"))
        p = torch.softmax(logits, dim=-1)                                                # fp32 inside, bf16 out
        y[:, :, i0:i1] = (p @ vv[:, :i1]).view(Hkv, G, B, D)
        pooled = p.view(Hkv, G, B, i1).mean(dim=1)                                       # model.py:413-418
        col_sum[:, :i1] += pooled.float().sum(dim=1)
        for w, acc in band.items():  # rows i of this block that still see column j inside a window of w
            inside = (torch.arange(i0, i1, device=q.device).view(-1, 1) - cols[:i1].view(1, -1)) < w
            acc[:, :i1] += (pooled.float() * inside).sum(dim=1)
        lo = max(i0, S - obs_len)
        if lo < i1:
            obs_rows[:, lo - (S - obs_len): i1 - (S - obs_len), :i1] = pooled[:, lo - i0:, :]
    stats = PrefillAttnStats(col_sum=col_sum.to(q.dtype).unsqueeze(0),
                             obs_mean=obs_rows.float().mean(dim=1).to(q.dtype).unsqueeze(0), obs_len=obs_len,
                             band_sum=band or None)
    return y.view(1, Hq, S, D), stats
