"""KV-cache policy layer: host-side mirror of the reference's cache.py surface.

Drop-in boundary: `get_cache_constructor(strategy) -> (cls, relevant_kwargs)`,
`cls(max_batch_size, n_heads, head_dim, dtype, **kwargs)`, `.update_kv / .update_state /
.return_attn / .reset / .compute_statistics`, `add_cache_arguments`, `cache_compatibility`
(reference cache.py:13-139, 152-346, 1444-1478).  Buffers carry the reference's names, shapes and
dtypes (cache.py:185-227, 573-574, 646-679) so `.to()`, `state_dict()`, `reset()` and code that
inspects them behave the same.

What differs is everything behind the surface: there is no per-op PyTorch implementation of the
decode step.  `decode_step()` runs eviction + overwrite + attention + score update as ONE CUDA
kernel launch (csrc/decode_step.cu) through `torch.ops.cold_compress_b200.decode_step`; the
three-call surface (`update_kv` -> attention -> `update_state`) maps onto the same kernel with a
phase mask.  CUDA only: on CPU tensors the ops raise.
"""
from __future__ import annotations

import argparse
import math
from collections import Counter

import torch
import torch.nn as nn

from . import _lib, ops
from .prompt_compression import get_prompt_compressor_constructor  # noqa: F401  (re-export, as the reference)

# --------------------------------------------------------------------------- CLI arguments

_STRATEGIES = ["full", "random", "recent_global", "heavy_hitter", "l2", "hybrid", "keep_it_odd"]

# (flag, kwargs, on_group) -- same flags, defaults and types as reference cache.py:13-118
_CACHE_FLAGS = [
    ("--max_cache_length", dict(type=float, default=[1.0], nargs="+",
                                help="Cache size per layer: 0<x<=1 is a fraction of |prompt|+max_new_tokens, >1 absolute. Tiled over layers."), True),
    ("--cache_bits", dict(default=None, type=int, choices=[2, 4, 8], help="Quantise the KV cache to this many bits."), True),
    ("--cache_length_pattern", dict(default="tile", choices=["tile", "repeat", "funnel", "pyramid"]), True),
    ("--cache_strategy", dict(default=["full"], nargs="+",
                              choices=_STRATEGIES + [f"debug_{s}" for s in _STRATEGIES]), True),
    ("--cache_strategy_pattern", dict(default="tile", choices=["tile", "repeat"],
                                      help="How cache_strategy is spread over layers."), True),
    ("--feed_long_prompts", dict(default=False, action="store_true",
                                 help="Prefill with prompt[:max_cache_length] and feed the rest token by token."), False),
    ("--prompt_compression_strategy", dict(default=["recent_global"], nargs="+",
                                           help="How a prompt longer than the cache is compressed into it."), True),
    ("--global_tokens", dict(default=1, type=int, help="Leading tokens that are never evicted."), True),
    ("--recent_window", dict(default=10, type=float, help="Most recent tokens spared from eviction (<1: fraction of cache)."), True),
    ("--history_window_size", dict(default=1, type=int, help="Heavy-hitter history window (1 = full history)."), True),
    ("--attn_thresholding", dict(default=False, action="store_true", help="Accumulate binary attention indicators."), True),
    ("--min_recovery_frac", dict(default=0.9, type=float, help="Hybrid (FastGen): minimum recovered attention mass."), False),
]


def add_cache_arguments(parser: argparse.ArgumentParser):
    group = parser.add_argument_group("cache_args")
    for flag, kw, on_group in _CACHE_FLAGS:
        (group if on_group else parser).add_argument(flag, **kw)


def cache_compatibility(args):
    """Same checks as reference cache.py:121-139."""
    triples = zip(args.max_cache_length, args.cache_strategy, args.prompt_compression_strategy)
    for length, strat, prompt_strat in triples:
        if strat == "heavy_hitter" and prompt_strat != "heavy_hitter":
            raise AssertionError("heavy_hitter caches need --prompt_compression_strategy heavy_hitter (it returns attention).")
        if strat == "hybrid" and getattr(args, "compile", False):
            raise AssertionError("hybrid is not supported with compile=True.")
        if strat in ("full", "hybrid") and length != 1.0:
            raise AssertionError(f"{strat} cache strategy only supports max_cache_length=1.0.")
    print("The cache argument values you provided appear compatible with each other!")


def create_window_attention_mask(seq_len, window_size, device, global_tokens: int = 4):
    """[seq_len, seq_len] bool: globals + a causal sliding window (reference cache.py:142-149), vectorised."""
    i = torch.arange(seq_len, device=device).view(-1, 1)
    j = torch.arange(seq_len, device=device).view(1, -1)
    return (j < global_tokens) | ((j <= i) & (j > i - window_size))


# --------------------------------------------------------------------------- base class


class KVCache(nn.Module):
    relevant_kwargs = ["max_cache_length", "global_tokens", "max_seq_length", "cache_bits"]
    policy = _lib.POLICY_EXTERNAL

    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, head_specific=False,
                 variable_length=False, **kwargs):
        super().__init__()
        for key, value in kwargs.items():
            setattr(self, key, value)
        if max_batch_size != 1:
            raise NotImplementedError("batch size is fixed to 1 (reference model.py:188-189)")
        self.cache_bits = getattr(self, "cache_bits", None)
        self.quantize, self.n_bit = self.cache_bits is not None, self.cache_bits
        if self.quantize and self.cache_bits not in (2, 4, 8):
            raise ValueError("Only 2-bit, 4-bit, and 8-bit quantization are supported")
        if not hasattr(self, "global_tokens"):
            self.global_tokens = 0
        self.n_heads, self.head_dim = n_heads, head_dim
        self.head_specific, self.variable_length = head_specific, variable_length
        L = self.max_cache_length
        self.cache_shape = (max_batch_size, n_heads, L, head_dim)
        self.register_buffer("k_cache", torch.zeros(self.cache_shape, dtype=dtype))
        self.register_buffer("v_cache", torch.zeros(self.cache_shape, dtype=dtype))
        self.register_buffer("pos", torch.full((max_batch_size, n_heads if head_specific else 1, L), -1, dtype=torch.int))
        self.register_buffer("cache_cts", torch.zeros(n_heads if variable_length else 1, dtype=torch.int))
        self.register_buffer("mask", torch.zeros((max_batch_size, n_heads, 1, L), dtype=torch.bool))
        if self.quantize:
            # --cache_bits: k_cache / v_cache above are the de-quantised WORKING copies the kernels read; after every
            # step they go through the reference's quantise -> de-quantise round trip (cc_requantize_cache), which
            # also leaves the packed codes and per-slot statistics the reference would hold (cache.py:176-200, 283-309)
            nbytes = n_heads * L * head_dim * self.n_bit // 8
            self.register_buffer("k_codes", torch.zeros(nbytes, dtype=torch.uint8))
            self.register_buffer("v_codes", torch.zeros(nbytes, dtype=torch.uint8))
            for name in ("k_scales", "v_scales", "k_zero_points", "v_zero_points"):
                self.register_buffer(name, torch.zeros(L, dtype=dtype))
        # kernel-side scratch (not part of the reference state)
        self.register_buffer("_sync_words", torch.zeros(8, dtype=torch.int32), persistent=False)
        self.register_buffer("_fill_idx", torch.zeros(n_heads, dtype=torch.int64), persistent=False)
        self.register_buffer("_empty", torch.zeros(0, dtype=torch.uint8), persistent=False)
        self._q_per_kv = None
        self._out = {}

    # ---- reference surface ------------------------------------------------------------
    def reset(self):
        self.k_cache.zero_()
        self.v_cache.zero_()
        self.mask.zero_()
        self.cache_cts.zero_()
        self.pos.fill_(-1)
        self._sync_words.zero_()

    def return_attn(self) -> bool:
        return False

    def memory_usage(self):
        tensors = [t for n, t in vars(self).items() if torch.is_tensor(t)]
        tensors += [b for n, b in self._buffers.items() if b is not None and not n.startswith("_")]
        return sum(t.element_size() * t.numel() for t in tensors) / (1024 ** 3)

    def compression_ratio(self, seq_len):
        n = seq_len - 1  # the final token is never passed to the cache
        assert torch.all(self.cache_cts <= self.max_cache_length)
        size = self.cache_cts.float()
        if self.n_bit is not None:
            size = size * (self.n_bit / 16.0)
        return ((n - size) / n).mean()

    def requantize(self):
        """End of a cache update with --cache_bits: the whole cache through quantise -> de-quantise (cache.py:338)."""
        if self.quantize:
            ops.requantize_cache(self.k_cache, self.v_cache, self.n_bit, self.k_codes, self.v_codes, self.k_scales,
                                 self.k_zero_points, self.v_scales, self.v_zero_points)

    def compute_statistics(self, seq_len):
        return {"compression_ratio": self.compression_ratio(seq_len).item(), "cache_memory_gb": self.memory_usage()}

    def return_kv_cache(self):
        return self.k_cache, self.v_cache, self.mask

    def update_kv(self, input_pos, k_val, v_val, is_prefill, **kwargs):
        """Reference cache.py:314-340.  Decode: eviction decision + in-place overwrite (kernel phase
        EVICT_FILL).  Prefill: contiguous fill of the (already compressed) prompt."""
        if is_prefill:
            self._prefill_update(input_pos, k_val, v_val, **kwargs)
        else:
            self._launch(_lib.PHASE_EVICT_FILL, None, k_val, v_val, input_pos, **kwargs)
        return self.return_kv_cache()

    def update_state(self, input_pos, k_val, v_val, is_prefill, attn, **kwargs):
        """No-op for caches that keep no score state (reference cache.py:342-346).  With --cache_bits the step ends
        here: the reference re-quantises at the end of update_kv but hands attention the un-quantised tensors, so in
        the three-call surface the round trip has to wait until attention has read the cache."""
        self.requantize()
        return None

    # ---- fused / phased decode step -----------------------------------------------------
    def attn_mode(self) -> int:
        return _lib.ATTN_NAIVE_BF16 if self.return_attn() else _lib.ATTN_MATH_FP32

    def decode_step(self, q, k_val, v_val, input_pos, **kwargs):
        """ONE kernel: evict -> overwrite -> attend -> pool -> score.  q [1,Hq,1,D]; returns y [1,Hq,1,D].
        overlap_predecessor=True promises that the kernel launched just before on this stream does not
        write this cache (CC_FLAG_OVERLAP_PREDECESSOR): the step then starts fetching early.
        warm_l2=[tensors]: read-only tensors (up to 3) the step pulls into L2 while HBM is idle."""
        y = self._launch(_lib.PHASE_ALL, q, k_val, v_val, input_pos, **kwargs)
        self.requantize()
        return y

    def attend(self, q, input_pos):
        """Attention over the current cache content only (kernel phase ATTEND): (y, pooled probs or None)."""
        y = self._launch(_lib.PHASE_ATTEND, q, None, None, input_pos)
        return y, (self._out["probs"].view(1, self.n_heads, 1, -1) if self.return_attn() else None)

    def _state_list(self):
        e = self._empty
        get = lambda n: getattr(self, n, None)  # noqa: E731
        names = ops.STATE_ORDER
        alias = {"sync_words": "_sync_words", "hyb_strategy_bits": "_hyb_bits", "hyb_recent_window": "_hyb_rw",
                 "hyb_heavy_budget": "_hyb_hh"}
        out = []
        for n in names:
            t = get(alias.get(n, n))
            out.append(t if torch.is_tensor(t) else e)
        return out

    def _meta(self, G):
        W = getattr(self, "history_window_size", 1)
        return [self.policy, self.attn_mode(), self.n_heads, G, self.head_dim, self.max_cache_length,
                int(self.global_tokens), int(getattr(self, "recent_window", 0) or 0), int(W),
                self.pos.shape[1], self.cache_cts.numel(), int(bool(getattr(self, "attn_thresholding", False)))]

    def _outputs(self, G, device, dtype):
        key = (G, device, dtype)
        if self._out.get("key") != key:
            self._out = {"key": key,
                         "y": torch.empty(self.n_heads * G, self.head_dim, device=device, dtype=dtype),
                         "probs": torch.empty(self.n_heads, self.max_cache_length, device=device, dtype=dtype)}
        return self._out

    def _external_fill_idx(self, input_pos):
        raise NotImplementedError

    def _launch(self, phases, q, k_val, v_val, input_pos, probs_in=None, **kwargs):
        dev = self.k_cache.device
        if dev.type != "cuda":
            raise _lib.CcError("the compressed-KV decode path is CUDA-only: move the model to a B200 (no CPU fallback)")
        if q is not None:
            G = q.shape[1] // self.n_heads
            self._q_per_kv = G
        else:
            G = self._q_per_kv or 1
        out = self._outputs(G, dev, self.k_cache.dtype)
        e = self._empty
        fill_idx = self._fill_idx  # static address: in for EXTERNAL, out (the eviction trace) otherwise
        if self.policy == _lib.POLICY_EXTERNAL and (phases & _lib.PHASE_EVICT_FILL):
            fill_idx.copy_(self._external_fill_idx(input_pos).to(torch.int64).reshape(-1).expand(self.n_heads))
        probs = out["probs"] if self.return_attn() else e
        if probs_in is not None:
            probs = probs_in.reshape(self.n_heads, -1).contiguous()
        ip = input_pos if input_pos.dtype == torch.int32 else input_pos.int()
        ops.decode_step(
            q.reshape(-1, self.head_dim) if q is not None else e,
            k_val.reshape(-1, self.head_dim) if k_val is not None else e,
            v_val.reshape(-1, self.head_dim) if v_val is not None else e,
            ip, self._state_list(), self._meta(G), out["y"], fill_idx, probs, phases,
            bool(kwargs.get("is_punc", False)),
            # a caller-chosen slot (EXTERNAL) is written to `fill_idx` by the kernel launched just before this one,
            # so the step must not start reading ahead of its predecessor
            (_lib.FLAG_OVERLAP_PREDECESSOR if kwargs.get("overlap_predecessor") and self.policy != _lib.POLICY_EXTERNAL else 0)
            | (_lib.FLAG_VIA_TOKEN_KERNEL if kwargs.get("via_token_kernel") else 0),
            list(kwargs.get("warm_l2", ())), kwargs.get("is_punc_dev"))
        return out["y"].view(1, self.n_heads * G, 1, self.head_dim) if q is not None else None

    # ---- prefill fill ---------------------------------------------------------------------
    def _prefill_update(self, input_pos, k_val, v_val, **kwargs):
        """Contiguous fill of slots [0, n) (reference cache.py:381-401)."""
        n = input_pos.shape[-1]
        self.pos[:, :, :n] = input_pos.int()
        self.k_cache[:, :, :n, :] = k_val
        self.v_cache[:, :, :n, :] = v_val
        if kwargs.get("update_mask", True):
            self.mask[:, :, :, :n] = True
        self.cache_cts += n if not self.variable_length else 0

    def fill_from_prompt(self, k_prompt, v_prompt, keep_idx):
        """Fused prompt-compression gather + fill: cache[:, j] = prompt[:, keep_idx[.., j]] (cc_prefill_gather)."""
        ops.prefill_gather(self._state_list(), self._meta(self._q_per_kv or 1), k_prompt.contiguous(),
                           v_prompt.contiguous(), keep_idx.to(torch.int64).contiguous())
        self.cache_cts += keep_idx.shape[-1] if not self.variable_length else 0


class KVCacheHeadConstant(KVCache):
    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, **kwargs):
        super().__init__(max_batch_size, n_heads, head_dim, dtype, head_specific=False, **kwargs)


class KVCacheHeadSpecific(KVCache):
    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, variable_length=False, **kwargs):
        super().__init__(max_batch_size, n_heads, head_dim, dtype, head_specific=True,
                         variable_length=variable_length, **kwargs)


# --------------------------------------------------------------------------- host-side (EXTERNAL) policies
# Their slot choice is a handful of integer ops on `pos`; the kernel takes the index and does the
# overwrite + attention.  Not in BASELINE configs.


def _base_eviction_idx(scores, pos_row, global_tokens):
    """reference cache.py:366-379."""
    scores = scores.clone().float().view(1, -1)
    scores[:, :global_tokens] = float("inf")
    scores.masked_fill_(pos_row.view(1, -1) == -1, float("-inf"))
    return torch.argmin(scores, dim=-1)


class KVCacheFull(KVCacheHeadConstant):
    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, **kwargs):
        kwargs = dict(kwargs)
        kwargs["global_tokens"] = 0  # every token is kept
        super().__init__(max_batch_size, n_heads, head_dim, dtype, **kwargs)

    def _external_fill_idx(self, input_pos):
        return self.pos[0, 0].argmin().view(1)  # first unfilled slot (reference cache.py:500-502)


class KVCacheRandom(KVCacheHeadConstant):
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "global_tokens", "recent_window"]

    def _external_fill_idx(self, input_pos):
        scores = torch.rand(self.max_cache_length, device=self.pos.device)
        scores[self.pos[0, 0] >= input_pos - self.recent_window] = float("inf")
        return _base_eviction_idx(scores, self.pos[0, 0], self.global_tokens)


class KVCacheKeepItOdd(KVCacheHeadConstant):
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "global_tokens", "recent_window"]

    def _external_fill_idx(self, input_pos):
        p = self.pos[0, 0]
        scores = (p % 2 == 1).float()
        scores[p >= input_pos - self.recent_window] = float("inf")
        return _base_eviction_idx(scores, p, self.global_tokens)


# --------------------------------------------------------------------------- in-kernel policies


class KVCacheRecentGlobal(KVCacheHeadConstant):
    """Sliding window + global tokens; the window is max_cache_length - global_tokens (reference cache.py:527-556)."""
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "global_tokens"]
    policy = _lib.POLICY_RECENT_GLOBAL


class KVCacheL2(KVCacheHeadSpecific):
    """Evicts the key with the largest L2 norm outside the recent window (reference cache.py:559-612)."""
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "global_tokens", "recent_window"]
    policy = _lib.POLICY_L2

    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, **kwargs):
        super().__init__(max_batch_size, n_heads, head_dim, dtype, **kwargs)
        if self.quantize:   # the reference takes the norm of its int8 cache at prefill and raises (cache.py:611)
            raise NotImplementedError("l2 cannot be combined with --cache_bits (nor can it in the reference)")
        self.register_buffer("key_norm", torch.zeros((max_batch_size, n_heads, self.max_cache_length), dtype=dtype))

    def reset(self):
        super().reset()
        self.key_norm.zero_()

    def update_state(self, input_pos, k_val, v_val, is_prefill, attn, **kwargs):
        if is_prefill:  # decode-time norms are written by the kernel together with the fill
            ops.key_norms(self.k_cache, self.key_norm)


class KVCacheHeavyHitter(KVCacheHeadSpecific):
    """Accumulated-attention eviction (ScissorHands / H2O; reference cache.py:615-765)."""
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "global_tokens", "history_window_size",
                       "recent_window", "attn_thresholding"]
    policy = _lib.POLICY_HEAVY_HITTER

    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, variable_length=False, **kwargs):
        super().__init__(max_batch_size, n_heads, head_dim, dtype, variable_length, **kwargs)
        W, L = self.history_window_size, self.max_cache_length
        thresh = bool(getattr(self, "attn_thresholding", False))
        if thresh and W == 1:
            # the reference's `bool_history[..., 0] += indicator` cannot be executed by torch (cache.py:719)
            raise ValueError("attn_thresholding needs history_window_size > 1")
        # bool indicators if thresholding, fp64 running sum if the whole history is kept, else a ring (cache.py:661-667)
        num_dtype = torch.bool if thresh else torch.float64 if W == 1 else dtype
        self.register_buffer("attn_history_num", torch.zeros((max_batch_size, n_heads, L, W), dtype=num_dtype))
        if W > 1 and self.policy == _lib.POLICY_HEAVY_HITTER:
            # exact running sum of each slot's ring, maintained by the kernel (the reference re-sums the ring)
            self.register_buffer("attn_history_sum", torch.zeros((max_batch_size, n_heads, L), dtype=torch.float64),
                                 persistent=False)
        self.register_buffer("attn_history_denom", torch.zeros((max_batch_size, n_heads, L), dtype=torch.int32))
        self.register_buffer("attn_counter", torch.zeros((1,), dtype=torch.int64))

    def reset(self):
        super().reset()
        self.attn_history_num.zero_()
        self.attn_history_denom.zero_()
        self.attn_counter.zero_()
        if hasattr(self, "attn_history_sum"):
            self.attn_history_sum.zero_()

    def resync_history(self):
        """Recompute the running ring sums from the ring (after the ring was written from outside the kernel)."""
        if hasattr(self, "attn_history_sum"):
            self.attn_history_sum.copy_(self.attn_history_num.to(torch.float64).sum(dim=-1))

    def return_attn(self) -> bool:
        return True

    def update_state(self, input_pos, k_val, v_val, is_prefill, attn, **kwargs):
        if not is_prefill:
            # kernel phase SCORE with the pooled probabilities the caller holds
            self._launch(_lib.PHASE_SCORE, None, None, None, input_pos, probs_in=attn.to(self.k_cache.dtype))
        else:
            self._seed_history(input_pos, attn)
        self.requantize()

    def _seed_history(self, input_pos, attn):
        """Prefill seed of the score history (reference cache.py:697-723): attn is either the full
        [1,H,S,S] map or the already-reduced cumulative attention [1,H,n]."""
        seq_len = attn.shape[-1]
        if attn.ndim == 4:
            attn = attn.squeeze(0).sum(dim=1) / (seq_len - input_pos)
        attn = attn.reshape(1, self.n_heads, -1, 1)
        if getattr(self, "attn_thresholding", False):
            attn = attn >= 1 / self.cache_cts      # indicator of above-uniform attention (cache.py:708)
        n = attn.shape[2]
        slot = int(self.attn_counter.item()) % self.history_window_size if self.history_window_size > 1 else 0
        col = self.attn_history_num[:, :, :, slot]
        if self.history_window_size == 1:
            col[:, :, :n] += attn[..., 0].to(col.dtype)
        else:
            col.zero_()
            col[:, :, :n] = attn[..., 0].to(col.dtype)
        self.attn_history_denom += 1
        self.attn_counter += 1
        self.resync_history()


class KVCacheHybrid(KVCacheHeavyHitter):
    """FastGen-style per-head hybrid strategies (reference cache.py:768-1288).  Decode runs in the
    fused kernel (policy HYBRID); head profiling at prefill is host-side tensor code."""
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "global_tokens", "token_ids",
                       "min_recovery_frac", "hybrid_strategies"]
    policy = _lib.POLICY_HYBRID

    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, **kwargs):
        self.attn_thresholding = False
        kwargs = dict(kwargs, history_window_size=400, recent_window=None, attn_thresholding=False)
        if kwargs.get("cache_bits") is not None:
            raise NotImplementedError("hybrid cannot be combined with --cache_bits")
        super().__init__(max_batch_size, n_heads, head_dim, dtype, variable_length=True, **kwargs)
        names = [s["strategy"] for s in self.hybrid_strategies]
        L = self.max_cache_length
        self.requires_special = any("special" in n for n in names)
        self.requires_punc = any("punc" in n for n in names)
        self.requires_heavy_hitter = any("heavy_hitter" in n for n in names)
        if self.requires_special:
            self.special_id_seqs = [list(ids) for ids in kwargs["token_ids"]["special"]]
            self.register_buffer("special_mask", torch.zeros((max_batch_size, n_heads, L), dtype=torch.bool))
            self.register_buffer("num_special", torch.zeros((1,), dtype=torch.int))
        if self.requires_punc:
            self.register_buffer("punc_ids", torch.tensor(kwargs["token_ids"]["punctuation"], dtype=torch.int64))
            self.register_buffer("punc_mask", torch.zeros((max_batch_size, n_heads, L), dtype=torch.bool))
            self.register_buffer("num_punc", torch.zeros((1,), dtype=torch.int))
        self.register_buffer("cache_strategies", torch.zeros(n_heads, dtype=torch.int64))
        # running fp64 sum of each slot's ring, maintained by the kernel (the reference re-sums the ring)
        self.register_buffer("attn_history_sum", torch.zeros((max_batch_size, n_heads, L), dtype=torch.float64),
                             persistent=False)
        self.register_buffer("_hyb_bits", torch.zeros(n_heads, dtype=torch.int32), persistent=False)
        self.register_buffer("_hyb_rw", torch.zeros(n_heads, dtype=torch.int32), persistent=False)
        self.register_buffer("_hyb_hh", torch.zeros(n_heads, dtype=torch.int32), persistent=False)

    def return_attn(self):
        return self.requires_heavy_hitter

    @staticmethod
    def strategy_bits(name: str) -> int:
        if name == "full":
            return _lib.HYB_FULL
        table = {"window": _lib.HYB_WINDOW, "heavy_hitter": _lib.HYB_HEAVY_HITTER, "special": _lib.HYB_SPECIAL,
                 "punc": _lib.HYB_PUNC}
        return sum(bit for key, bit in table.items() if key in name)

    def set_head_strategies(self, strategy_ids):
        """Install the per-head strategy choice (reference: self.cache_strategies, cache.py:1200)."""
        ids = torch.as_tensor(strategy_ids, dtype=torch.int64)
        L = self.max_cache_length
        bits = [self.strategy_bits(self.hybrid_strategies[int(i)]["strategy"]) for i in ids]
        rw = [round(self.hybrid_strategies[int(i)].get("recent_window", 0) * L) for i in ids]
        hh = [round(self.hybrid_strategies[int(i)].get("heavy_hitter_frac", 0) * L) for i in ids]
        dev = self.cache_strategies.device
        self.cache_strategies.copy_(ids.to(dev))
        self._hyb_bits.copy_(torch.tensor(bits, dtype=torch.int32))
        self._hyb_rw.copy_(torch.tensor(rw, dtype=torch.int32))
        self._hyb_hh.copy_(torch.tensor(hh, dtype=torch.int32))
        names = [self.hybrid_strategies[int(i)]["strategy"] for i in ids]
        self.requires_heavy_hitter = any("heavy_hitter" in n for n in names)

    def reset(self):
        super().reset()
        names = [s["strategy"] for s in self.hybrid_strategies]
        self.requires_heavy_hitter = any("heavy_hitter" in n for n in names)
        for n in ("special_mask", "num_special", "punc_mask", "num_punc"):
            if hasattr(self, n):
                getattr(self, n).zero_()

    def _launch(self, phases, q, k_val, v_val, input_pos, probs_in=None, **kwargs):
        if "is_punc" not in kwargs and hasattr(self, "punc_ids") and kwargs.get("input_ids") is not None:
            # device-side flag: no host read, so the step stays capturable in a CUDA graph
            # (plain compare + any: torch.isin is not capturable)
            tok = kwargs["input_ids"].reshape(-1)[-1:].to(self.punc_ids.dtype)
            kwargs["is_punc_dev"] = (tok.view(1, 1) == self.punc_ids.view(1, -1)).any().to(torch.uint8).view(1)
        return super()._launch(phases, q, k_val, v_val, input_pos, probs_in=probs_in, **kwargs)

    # ---- prefill: head profiling (reference cache.py:1066-1272) --------------------------------------
    def profiling_windows(self, seq_len):
        """Window sizes (in tokens) whose band sums the prefill attention must provide."""
        out = set()
        for s in self.hybrid_strategies:
            if "window" in s["strategy"]:
                out.add(max(1, int(s["recent_window"] * seq_len)))
        return sorted(out)

    def _special_mask(self, input_ids):
        ids = input_ids.reshape(-1)
        m = torch.zeros_like(ids, dtype=torch.bool)
        for seq in self.special_id_seqs:
            n = len(seq)
            if n == 1:
                m |= ids == seq[0]
            else:
                t = torch.tensor(seq, device=ids.device, dtype=ids.dtype)
                win = ids.unfold(0, n, 1) if ids.numel() >= n else ids.new_zeros((0, n))
                hit = (win == t).all(dim=1).nonzero().flatten()
                for k in range(n):
                    m[hit + k] = True
        return m

    def _column_masks(self, cum_attn, special, punc, total_len, seq_len):
        """Per strategy: (column mask [H, S] of tokens kept for every query, window size or 0).
        Same construction as reference build_masks (cache.py:1066-1136), without the S x S tensors:
        a strategy's mask is `column mask | causal band of `window` tokens`."""
        H = cum_attn.shape[0] if cum_attn is not None else self.n_heads
        dev = self.pos.device
        cols = torch.arange(seq_len, device=dev)
        out = []
        for s in self.hybrid_strategies:
            name = s["strategy"]
            col = (cols < self.global_tokens).expand(H, -1).clone()
            if "special" in name:
                col |= special.view(1, -1)
            if "punc" in name:
                col |= punc.view(1, -1)
            window = max(1, int(s["recent_window"] * total_len)) if "window" in name else 0
            if "heavy_hitter" in name:
                # candidates: tokens the LAST query does not already keep (cache.py:1103)
                last_row = col[0] | (cols > seq_len - 1 - window) if window else col[0]
                avail = (~last_row).nonzero().flatten()
                num_hh = math.ceil(min(s["heavy_hitter_frac"] * total_len, avail.numel()))
                if num_hh > 0:
                    top = cum_attn[:, avail].topk(num_hh, dim=1).indices
                    col.scatter_(1, avail[top], True)
            if name == "full":
                col.fill_(True)
            out.append((col, window))
        return out

    @torch.no_grad()
    def profile_and_update(self, input_pos, k_val, v_val, attn, **kwargs):
        """Choose a strategy per head from the prompt's attention (FastGen), lay the prompt out with the
        kept tokens first, seed masks / counts / history.  `attn` is a PrefillAttnStats with band sums
        (or the reference's full [1,H,S,S] map, reduced here).  Deviations from the reference, both
        documented in DESIGN.md: recovered mass is computed from column and band sums (same value up to
        bf16 rounding of the reference's row sums), and kept tokens are laid out in ascending position
        (the reference's order comes from an unstable argsort)."""
        from .prompt_compression import PrefillAttnStats
        input_ids = kwargs["input_ids"].reshape(-1)
        S, H, dev = input_ids.numel(), self.n_heads, self.pos.device
        assert self.max_cache_length >= S
        special = self._special_mask(input_ids) if self.requires_special else torch.zeros(S, dtype=torch.bool, device=dev)
        punc = torch.isin(input_ids, self.punc_ids) if hasattr(self, "punc_ids") else torch.zeros(S, dtype=torch.bool, device=dev)
        if hasattr(self, "num_special"):
            self.num_special.fill_(int(special.sum()))
        if hasattr(self, "num_punc"):
            self.num_punc.fill_(int(punc.sum()))
        if not isinstance(attn, PrefillAttnStats):  # full map given: reduce it
            full = attn[0]
            wins = self.profiling_windows(S)
            idx = torch.arange(S, device=dev)
            band = {w: (full.float() * ((idx.view(-1, 1) - idx.view(1, -1)) < w)).sum(dim=1) for w in wins}
            attn = PrefillAttnStats(col_sum=full.sum(dim=1).unsqueeze(0), obs_mean=None, obs_len=0, band_sum=band)
        col_sum = attn.col_sum[0].float()                        # [H, S]
        cum_attn = (attn.col_sum[0] / (S - input_pos)).to(self.k_cache.dtype)   # cache.py:1158
        # 1. recovered attention mass of every strategy (cache.py:1165-1173)
        scores = []
        for col, window in self._column_masks(cum_attn.float(), special, punc, S, S):
            mass = (col_sum * col).sum(dim=1)
            if window:
                mass = mass + (attn.band_sum[window] * (~col)).sum(dim=1)
            scores.append(mass / S)                              # mean over the S queries (rows sum to 1)
        scores = torch.stack(scores)                             # [n_strategies, H]
        strategies = (scores >= self.min_recovery_frac).int().argmax(dim=0)
        self.set_head_strategies(strategies.cpu())
        # 2. tokens each head keeps, with budgets relative to the cache length (cache.py:1177-1185)
        fill = self._column_masks(cum_attn.float(), special, punc, self.max_cache_length, S)
        cols = torch.arange(S, device=dev)
        keep = torch.zeros(H, S, dtype=torch.bool, device=dev)
        for h in range(H):
            col, window = fill[int(strategies[h])]
            keep[h] = col[h] | ((cols > S - 1 - window) if window else False)
        # 3. kept tokens first (ascending), then the rest; everything is copied, unkept slots are cleared
        order = torch.sort((~keep).int() * S + cols.view(1, -1), dim=1).indices
        ex = order.view(1, H, S, 1).expand(-1, -1, -1, self.head_dim)
        self.k_cache[:, :, :S] = k_val.gather(2, ex)
        self.v_cache[:, :, :S] = v_val.gather(2, ex)
        self.pos[:, :, :S] = input_pos.view(1, -1).expand(H, -1).gather(1, order).int()
        cts = keep.sum(dim=1).int()
        self.cache_cts.copy_(cts)
        live = cols.view(1, -1) < cts.view(-1, 1)
        self.pos[0, :, :S].masked_fill_(~live, -1)
        self.k_cache[0, :, :S].masked_fill_(~live.unsqueeze(-1), 0)
        self.v_cache[0, :, :S].masked_fill_(~live.unsqueeze(-1), 0)
        self.mask[0, :, 0, :S] = live
        if hasattr(self, "special_mask"):
            self.special_mask[0, :, :S] = special.view(1, -1).expand(H, -1).gather(1, order)
        if hasattr(self, "punc_mask"):
            self.punc_mask[0, :, :S] = punc.view(1, -1).expand(H, -1).gather(1, order)
        if self.requires_heavy_hitter:
            self._seed_history(input_pos, cum_attn.gather(1, order).unsqueeze(0))

    def update_state(self, input_pos, k_val, v_val, is_prefill, attn, **kwargs):
        if is_prefill:
            return self.profile_and_update(input_pos, k_val, v_val, attn, **kwargs)
        if self.requires_heavy_hitter:
            super().update_state(input_pos, k_val, v_val, is_prefill, attn, **kwargs)

    def compute_statistics(self, seq_len):
        stats = super().compute_statistics(seq_len)
        ids = self.cache_strategies.tolist()
        cts = Counter(self.hybrid_strategies[i]["strategy"] for i in ids)
        stats["avg_strategy_idx"] = sum(ids) / len(ids)
        for name in sorted({s["strategy"] for s in self.hybrid_strategies}):
            stats[name] = cts.get(name, 0) / len(ids)
        return stats



class KVCacheAnalysis(KVCacheFull):
    """`debug_<strategy>`: a full cache that also drives a shadow cache with the compression strategy and records the
    attention mass the strategy would have lost at every decode step (reference cache.py:1291-1420).

    Decode = the same kernel three times: the full cache appends and attends (returning the pooled probabilities),
    the shadow cache takes its eviction decision + overwrite (phase EVICT_FILL) and its score update (phase SCORE)
    from the full cache's probabilities gathered at the positions the shadow still holds."""
    relevant_kwargs = ["max_cache_length", "max_seq_length", "cache_bits", "history_window_size", "recent_window",
                       "attn_thresholding", "global_tokens", "prompt_compression_strategy"]

    def __init__(self, max_batch_size, n_heads, head_dim, dtype=torch.bfloat16, cache_strategy="heavy_hitter", **kwargs):
        full_kwargs = {"global_tokens": 0, "max_cache_length": kwargs["max_seq_length"],
                       "prompt_compression_strategy": kwargs["prompt_compression_strategy"]}
        super().__init__(max_batch_size, n_heads, head_dim, dtype, **full_kwargs)
        self.compressed = get_cache_constructor(cache_strategy)[0](max_batch_size, n_heads, head_dim, dtype, **kwargs)
        if not self.compressed.return_attn():
            # the reference's update_state dereferences attn (None for strategies that return no attention)
            raise ValueError(f"debug_{cache_strategy}: the attention-loss analysis needs a strategy that returns attention")
        self.register_buffer("attention_losses", torch.full((self.max_cache_length,), -1, dtype=dtype))
        self.register_buffer("attention_loss_ctr", torch.zeros((1,), dtype=torch.int))
        self.prompt_compressor = get_prompt_compressor_constructor(self.prompt_compression_strategy)(
            head_specific=self.compressed.head_specific, **kwargs)
        self.head_specific = self.compressed.head_specific

    def return_attn(self):
        return self.compressed.return_attn()

    def reset(self):
        super().reset()
        self.compressed.reset()
        self.attention_losses.fill_(-1)
        self.attention_loss_ctr.zero_()

    def memory_usage(self):
        return super().memory_usage() + self.compressed.memory_usage()

    # ---- reference surface -----------------------------------------------------------------
    def update_kv(self, input_pos, k_val, v_val, is_prefill, **kwargs):
        out = super().update_kv(input_pos, k_val, v_val, is_prefill, **kwargs)
        if input_pos.shape[-1] < self.compressed.max_cache_length:   # else: compressed in update_state (needs attention)
            self.compressed.update_kv(input_pos, k_val, v_val, is_prefill, **kwargs)
        return out

    def _cum_attn(self, input_pos, attn):
        from .prompt_compression import PrefillAttnStats
        if isinstance(attn, PrefillAttnStats):
            return attn.col_sum / (attn.col_sum.shape[-1] - input_pos)
        return attn

    def update_state(self, input_pos, k_val, v_val, is_prefill, attn, **kwargs):
        comp = self.compressed
        if is_prefill and input_pos.shape[-1] > comp.max_cache_length:
            keep, k_c, v_c, cum = self.prompt_compressor(input_pos, k_val, v_val, attn=attn)
            comp.update_kv(keep, k_c, v_c, True)
            comp.update_state(keep, k_c, v_c, True, cum)
        elif is_prefill:
            comp.update_state(input_pos, k_val, v_val, True, self._cum_attn(input_pos, attn))
        else:
            self._shadow_score(input_pos, k_val, v_val, attn)

    def _shadow_score(self, input_pos, k_val, v_val, attn):
        """attn [1, H, (1,) Lfull]: pooled probabilities over the full cache (slot == position)."""
        comp = self.compressed
        attn = attn.reshape(1, self.n_heads, -1)
        idx = comp.pos.long().expand(1, self.n_heads, -1)
        idx = idx.masked_fill(idx < 0, attn.shape[-1] - 1)   # unfilled shadow slots read the last (still empty, 0.0) position
        attn_c = attn.gather(2, idx)
        comp.update_state(input_pos, k_val, v_val, False, attn_c)
        loss = (1 - attn_c.sum(dim=-1)).mean()         # attention mass of the evicted tokens
        self.attention_losses.index_put_((self.attention_loss_ctr.long(),), loss.to(self.attention_losses.dtype))
        self.attention_loss_ctr += 1

    # ---- decode: full-cache step, then the shadow's two phases ---------------------------------
    def decode_step(self, q, k_val, v_val, input_pos, **kwargs):
        kw = {k: v for k, v in kwargs.items() if k in ("overlap_predecessor", "input_ids")}
        y = self._launch(_lib.PHASE_EVICT_FILL | _lib.PHASE_ATTEND, q, k_val, v_val, input_pos, **kw)
        self.compressed._q_per_kv = self._q_per_kv
        self.compressed.update_kv(input_pos, k_val, v_val, False, input_ids=kwargs.get("input_ids"))
        self._shadow_score(input_pos, k_val, v_val, self._out["probs"])
        return y

    def compute_statistics(self, seq_len):
        stats = super().compute_statistics(seq_len)
        losses = self.attention_losses[: int(self.attention_loss_ctr)]
        assert not torch.any(losses == -1)
        for k in range(500, len(losses), 500):
            stats[f"attention_loss@{k}"] = losses[:k].mean().item()
        stats["attention_loss"] = losses.mean().item()
        return stats


# --------------------------------------------------------------------------- registry

_REGISTRY = {
    "full": KVCacheFull, "l2": KVCacheL2, "random": KVCacheRandom, "recent_global": KVCacheRecentGlobal,
    "heavy_hitter": KVCacheHeavyHitter, "hybrid": KVCacheHybrid, "keep_it_odd": KVCacheKeepItOdd,
}


def get_cache_constructor(cache_strategy):
    """(constructor, relevant kwargs) for a strategy name (reference cache.py:1444-1478)."""
    if cache_strategy in _REGISTRY:
        cls = _REGISTRY[cache_strategy]
        return cls, cls.relevant_kwargs
    if cache_strategy.startswith("debug"):
        import re
        inner = re.sub(r"debug_+", "", cache_strategy).strip()
        relevant = get_cache_constructor(inner)[1] + ["prompt_compression_strategy"]

        def make(max_batch_size, n_heads, head_dim, dtype, **kwargs):
            return KVCacheAnalysis(max_batch_size, n_heads, head_dim, dtype, cache_strategy=inner, **kwargs)
        return make, relevant
    raise ValueError(f"Invalid cache strategy: {cache_strategy}")
