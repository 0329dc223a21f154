"""Prompt compressors: pick the `max_cache_length` prompt tokens that enter the cache when the
prompt is longer than the cache (reference prompt_compression.py).

Surface kept: `get_prompt_compressor_constructor(strategy) -> cls`, `cls(head_specific, **kwargs)`,
`compressor(input_pos, k_val, v_val, attn=...) -> (keep_idxs, k_val, v_val, attn_or_None)`.

`attn` may be the reference's full pooled attention map `[1, Hkv, S, S]`, or a `PrefillAttnStats`
holding only the two reductions the compressors and caches consume (column sums and the mean of
the last `observation_len` rows) -- which is what the prefill path produces so that the S x S map
is never materialised.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class PrefillAttnStats:
    """Reductions of the pooled causal attention map P [Hkv, S, S] (bf16 semantics of the reference):
    col_sum[h, j] = bf16(sum_i P[h, i, j])  (prompt_compression.py:192, cache.py:704)
    obs_mean[h, j] = bf16(mean over the last `obs_len` rows i of P[h, i, j])  (prompt_compression.py:175)."""
    col_sum: torch.Tensor   # [1, Hkv, S] bf16
    obs_mean: torch.Tensor  # [1, Hkv, S] bf16
    obs_len: int
    # optional, for hybrid head profiling (cache.py:1066-1187): for each requested window size W,
    # band_sum[W][h, j] = sum over the W queries i in [j, j+W) of P[h, i, j] (fp32), i.e. the attention a
    # sliding window of W tokens recovers at column j
    band_sum: dict = None

    @property
    def shape(self):
        s = self.col_sum.shape[-1]
        return (1, self.col_sum.shape[1], s, s)

    @property
    def ndim(self):
        return 4


class PromptCompressor:
    """priority -> top-k -> ascending order -> gather (reference prompt_compression.py:21-43)."""
    head_specific_only = False

    def __init__(self, head_specific, **kwargs):
        for key, value in kwargs.items():
            setattr(self, key, value)
        self.head_specific = head_specific
        assert self.is_compatible(), (
            f"Prompt compressor ({type(self).__name__}) is not compatible with the chosen cache strategy.")

    def is_compatible(self) -> bool:
        return self.head_specific or not self.head_specific_only

    def _recent_global_mask(self, input_pos):
        seq_len = input_pos.shape[-1]
        return (input_pos < self.global_tokens) | (input_pos >= seq_len - self.recent_window)

    def _keep_idxs(self, priority):
        top = priority.topk(self.max_cache_length, dim=-1).indices
        return top.sort(dim=-1).values.squeeze(0)

    def _token_importances(self, input_pos, k_val, v_val, **kwargs):
        raise NotImplementedError

    def _filter_kv(self, keep_idxs, k_val, v_val):
        if keep_idxs.dim() == 1:  # same tokens for every head
            return k_val[:, :, keep_idxs], v_val[:, :, keep_idxs]
        idx = keep_idxs.view(1, -1, self.max_cache_length, 1).expand(-1, -1, -1, k_val.shape[-1])
        return k_val.gather(2, idx), v_val.gather(2, idx)

    def _update_state(self, keep_idxs, input_pos, **kwargs):
        return None

    def select(self, input_pos, k_val, v_val, **kwargs):
        """keep_idxs only (the fused path gathers straight into the cache with cc_prefill_gather)."""
        return self._keep_idxs(self._token_importances(input_pos, k_val, v_val, **kwargs))

    def __call__(self, input_pos, k_val, v_val, **kwargs):
        keep_idxs = self.select(input_pos, k_val, v_val, **kwargs)
        k_val, v_val = self._filter_kv(keep_idxs, k_val, v_val)
        return keep_idxs, k_val, v_val, self._update_state(keep_idxs, input_pos, **kwargs)


class PromptCompressorFull(PromptCompressor):
    """Pass-through (reference prompt_compression.py:91-106)."""

    def __call__(self, input_pos, k_val, v_val, **kwargs):
        return input_pos, k_val, v_val, None

    def select(self, input_pos, k_val, v_val, **kwargs):
        return input_pos


class PromptCompressorRandom(PromptCompressor):
    def _token_importances(self, input_pos, k_val, v_val, **kwargs):
        seq_len = input_pos.shape[-1]
        save = self._recent_global_mask(input_pos)
        priority = torch.where(save, torch.full_like(input_pos, seq_len), torch.full_like(input_pos, -seq_len))
        return priority + torch.randperm(seq_len, device=priority.device)


class PromptCompressorRecentGlobal(PromptCompressor):
    def __init__(self, head_specific, **kwargs):
        super().__init__(head_specific, **kwargs)
        assert self.max_cache_length - self.global_tokens > 0, (
            f"Number of global tokens ({self.global_tokens}) cannot exceed the max cache length ({self.max_cache_length})")

    def _token_importances(self, input_pos, k_val, v_val, **kwargs):
        # newest positions first, global tokens always (reference prompt_compression.py:141-145)
        return input_pos.masked_fill(input_pos < self.global_tokens, input_pos.shape[-1])


class PromptCompressorHeavyHitter(PromptCompressor):
    """SnapKV: attention of the last `observation_len` queries, smoothed by a 5-tap average
    (reference prompt_compression.py:148-194)."""
    head_specific_only = True
    kernel_size = 5
    observation_len = 16

    def _token_importances(self, input_pos, k_val, v_val, **kwargs):
        attn = kwargs["attn"]
        seq_len = input_pos.shape[-1]
        obs_len = min(self.observation_len, seq_len)
        if isinstance(attn, PrefillAttnStats):
            assert attn.obs_len == obs_len
            priority = attn.obs_mean.clone()
        else:
            priority = attn[:, :, -obs_len:, :].mean(dim=2)
        priority = F.avg_pool1d(priority, self.kernel_size, stride=1, padding=self.kernel_size // 2,
                                ceil_mode=False, count_include_pad=False)
        priority[:, :, -obs_len:] = 1.0
        priority[:, :, : self.global_tokens] = 1.0
        return priority

    def _update_state(self, keep_idxs, input_pos, **kwargs):
        attn = kwargs["attn"]
        seq_len = input_pos.shape[-1]
        col_sum = attn.col_sum if isinstance(attn, PrefillAttnStats) else attn.sum(dim=2)
        cum_attn = col_sum / (seq_len - input_pos)
        return cum_attn.gather(2, keep_idxs.view(1, -1, self.max_cache_length))


class PromptCompressorL2(PromptCompressor):
    head_specific_only = True

    def _token_importances(self, input_pos, k_val, v_val, **kwargs):
        priority = -torch.linalg.vector_norm(k_val, ord=2, dim=-1)
        return priority.masked_fill(self._recent_global_mask(input_pos).view(1, 1, -1), float("inf"))


class PromptCompressorKeepItOdd(PromptCompressor):
    def _token_importances(self, input_pos, k_val, v_val, **kwargs):
        seq_len = input_pos.shape[-1]
        priority = input_pos.masked_fill(self._recent_global_mask(input_pos), seq_len * 2)
        priority[input_pos % 2 == 0] -= seq_len
        return priority


_COMPRESSORS = {
    "full": PromptCompressorFull, "recent_global": PromptCompressorRecentGlobal,
    "heavy_hitter": PromptCompressorHeavyHitter, "l2": PromptCompressorL2,
    "random": PromptCompressorRandom, "keep_it_odd": PromptCompressorKeepItOdd,
}


def get_prompt_compressor_constructor(strategy):
    try:
        return _COMPRESSORS[strategy]
    except KeyError:
        raise ValueError(f"Unknown prompt compression strategy: {strategy}") from None
