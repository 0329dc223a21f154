"""CPU restatement of the reference's prefill attention + pooled-map reductions.  TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module (the product path is csrc/prefill_attn.cu through
cold_compress_b200.attention_utils.prefill_attention_stats, CUDA only).

* `full_map(q, k, v)` follows the reference op for op through the materialised map:
  attention_utils.py:36-54 (q @ k^T * scale -> + mask bias -> softmax -> @ v, every op in bf16 with
  fp32 accumulation inside) and model.py:413-418 (GQA mean).  Small S only.
* `reductions(pooled, obs_len, windows)` are the consumers' reductions: prompt_compression.py:175
  (observation-window mean), :192 and cache.py:704 (column sums), cache.py:1066-1173 (window bands).
* `blockwise_stats(q, k, v, ...)` is the same computation block by block (no S x S tensor), for
  host-logic tests at sizes where the full map is too large.
Pinned against outputs of the unmodified reference in tests/golden/prefill_*.npz (tests/test_prefill_oracle.py).
"""
from __future__ import annotations

import math

import torch

from cold_compress_b200.prompt_compression import PrefillAttnStats


@torch.no_grad()
def full_map(q, k, v):
    """q [1,Hq,S,D], k/v [1,Hkv,S,D] bf16 on CPU -> (y [1,Hq,S,D] bf16, pooled [1,Hkv,S,S] bf16)."""
    Hq, Hkv = q.shape[1], k.shape[1]
    G, S, D = Hq // Hkv, q.shape[2], q.shape[3]
    kk = k.repeat_interleave(G, dim=1)          # model.py:395-400
    vv = v.repeat_interleave(G, dim=1)
    w = q @ kk.transpose(-2, -1) * (1.0 / math.sqrt(D))
    mask = torch.tril(torch.ones(S, S, dtype=torch.bool))
    w = w + torch.zeros_like(w).masked_fill_(mask.logical_not(), float("-inf"))
    p = torch.softmax(w, dim=-1)
    y = p @ vv
    pooled = p.view(1, Hkv, G, S, S).mean(dim=2)
    return y, pooled


@torch.no_grad()
def reductions(pooled, obs_len=16, windows=()):
    S = pooled.shape[-1]
    obs_len = min(obs_len, S)
    idx = torch.arange(S)
    band = {int(w): (pooled[0].float() * ((idx.view(-1, 1) - idx.view(1, -1)) < w)).sum(dim=1) for w in windows}
    return PrefillAttnStats(col_sum=pooled.sum(dim=2), obs_mean=pooled[:, :, -obs_len:, :].mean(dim=2),
                            obs_len=obs_len, band_sum=band or None)


@torch.no_grad()
def blockwise_stats(q, k, v, obs_len: int = 16, block: int = 512, windows=()):
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
