"""Attention wrappers around the cache (reference attention_utils.py:8-54).

* `scaled_dot_product_attention` keeps the reference signature for callers that hold explicit
  K/V tensors (prefill, or generic use).  It is plain tensor code: the decode step never comes
  through here -- `Attention.forward` hands decode to `KVCache.decode_step` (one CUDA kernel).
* `prefill_attention_stats` is the prefill path for caches that need attention probabilities:
  one tcgen05 kernel (csrc/prefill_attn.cu) produces y plus the reductions of the pooled
  probability map that the prompt compressors and caches consume (`PrefillAttnStats`), so the
  [Hq, S, S] map the reference materialises (4 GiB at S=8192) never exists.  Rounding points
  follow the reference's naive path: bf16 after q@k^T, after *scale, after softmax, after the
  GQA mean.  `prefill_attention` is the same kernel without statistics.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch.nn import functional as F

from . import _lib, ops
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


_workspace: dict = {}


def _scratch(device, nbytes: int) -> torch.Tensor:
    """Per-device scratch for the partial column sums (grown on demand, reused across layers)."""
    buf = _workspace.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspace[device] = buf
    return buf


@torch.no_grad()
def prefill_attention(q, k, v, attn_mode: int = _lib.ATTN_MATH_FP32):
    """Causal GQA attention over a prompt on the tensor cores, no statistics (caches that keep no scores).
    q [1, Hq, S, D], k/v [1, Hkv, S, D] bf16 (any row/head strides).  Returns y [1, Hq, S, D] (a view of [S, Hq, D])."""
    _, Hq, S, D = q.shape
    y = torch.empty(S, Hq, D, dtype=q.dtype, device=q.device)
    ops.prefill_attention(q[0], k[0], v[0], y.permute(1, 0, 2), attn_mode, 0, [], None, None, None, None)
    return y.permute(1, 0, 2).unsqueeze(0)


@torch.no_grad()
def prefill_attention_stats(q, k, v, obs_len: int = 16, windows=()):
    """Causal self-attention over a prompt with GQA, plus the reductions of the pooled probability map.

    q [1, Hq, S, D], k/v [1, Hkv, S, D] (bf16, post-RoPE; any row/head strides).  Returns
    (y [1, Hq, S, D], PrefillAttnStats).  ONE tcgen05 kernel (csrc/prefill_attn.cu) + a fixed-order reduction
    of its per-tile partial sums; the reference's rounding points (naive path) are applied in the kernel,
    the final bf16 roundings of the reductions (prompt_compression.py:175, 192) here."""
    if not q.is_cuda:
        raise _lib.CcError("prefill_attention_stats is CUDA-only (cc_prefill_attention); there is no CPU path")
    _, Hq, S, D = q.shape
    Hkv = k.shape[1]
    G = Hq // Hkv
    obs_len = min(obs_len, S)
    windows = [int(w) for w in windows]
    dev = q.device
    y = torch.empty(S, Hq, D, dtype=q.dtype, device=dev)
    col = torch.empty(Hkv, S, dtype=torch.float32, device=dev)
    obs = torch.empty(Hkv, S, dtype=torch.float32, device=dev)
    band = torch.empty(len(windows), Hkv, S, dtype=torch.float32, device=dev) if windows else None
    ws = _scratch(dev, ops.prefill_attention_workspace(S, Hkv, G, obs_len, len(windows)))
    ops.prefill_attention(q[0], k[0], v[0], y.permute(1, 0, 2), _lib.ATTN_NAIVE_BF16, obs_len, windows, col, obs, band, ws)
    stats = PrefillAttnStats(col_sum=col.to(q.dtype).unsqueeze(0),
                             obs_mean=(obs / obs_len).to(q.dtype).unsqueeze(0), obs_len=obs_len,
                             band_sum={w: band[i] for i, w in enumerate(windows)} if windows else None)
    return y.permute(1, 0, 2).unsqueeze(0), stats
