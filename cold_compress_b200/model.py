"""Model shell that hosts the compressed-KV attention layer (reference model.py).

Kept: class / attribute / parameter names (`tok_embeddings`, `layers.N.attention.wqkv`, `wo`,
`feed_forward.w1/w2/w3`, `attention_norm`, `ffn_norm`, `norm`, `output`) so reference checkpoints
load, `Transformer.setup_caches / forward / reset_caches / get_cache_stats`, and the
`Attention.forward(x, input_ids, freqs_cis, mask, is_prefill, input_pos, attn_top_k)` signature
(model.py:363-372).

Different: the decode branch of `Attention.forward` is ONE fused CUDA kernel per layer
(`kv_cache.decode_step`), K/V are never `repeat_interleave`d, and prefill for
probability-consuming caches goes through `prefill_attention_stats` instead of an S x S map.
"""
from __future__ import annotations

import math
import os
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor
from torch.nn import functional as F

from . import ops
from .attention_utils import prefill_attention_stats, scaled_dot_product_attention
from .cache import get_cache_constructor
from .prompt_compression import PrefillAttnStats, get_prompt_compressor_constructor


def find_multiple(n: int, k: int) -> int:
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class ModelArgs:
    block_size: int = 2048
    vocab_size: int = 32000
    n_layer: int = 32
    n_head: int = 32
    dim: int = 4096
    intermediate_size: int = None
    n_local_heads: int = -1
    head_dim: int = 64
    rope_base: float = 10000
    norm_eps: float = 1e-5
    attention_bias: bool = False
    max_length: int = 4096
    rope_scaling: Optional[Dict[str, Any]] = None

    def __post_init__(self):
        if self.n_local_heads == -1:
            self.n_local_heads = self.n_head
        if self.intermediate_size is None:
            self.intermediate_size = find_multiple(int(2 * (4 * self.dim) / 3), 256)
        self.head_dim = self.dim // self.n_head

    @classmethod
    def from_name(cls, name: str):
        if name in transformer_configs:
            return cls(**transformer_configs[name])
        hits = sorted((c for c in transformer_configs if c in str(name).upper() or c in str(name)), key=len,
                      reverse=True)
        assert hits, f"unknown model name {name}"
        assert len(hits) == 1 or len(hits[0]) != len(hits[1]), name
        return cls(**transformer_configs[hits[0]])


def _llama3(block, **extra):
    return dict(block_size=block, n_layer=32, n_head=32, n_local_heads=8, dim=4096, intermediate_size=14336,
                vocab_size=128256, rope_base=500000, max_length=block, **extra)


def _qwen2(n_layer, n_head, n_kv, dim, inter, vocab):
    return dict(block_size=32768, n_layer=n_layer, n_head=n_head, n_local_heads=n_kv, dim=dim, intermediate_size=inter,
                vocab_size=vocab, rope_base=1000000, attention_bias=True, norm_eps=1e-6, max_length=32768)


# same model zoo as the reference (model.py:74-171)
transformer_configs = {
    "CodeLlama-7b-Python-hf": dict(block_size=16384, vocab_size=32000, n_layer=32, dim=4096, rope_base=1000000),
    "7B": dict(n_layer=32, n_head=32, dim=4096),
    "13B": dict(n_layer=40, n_head=40, dim=5120),
    "30B": dict(n_layer=60, n_head=52, dim=6656),
    "34B": dict(n_layer=48, n_head=64, dim=8192, vocab_size=32000, n_local_heads=8, intermediate_size=22016,
                rope_base=1000000),
    "70B": dict(n_layer=80, n_head=64, dim=8192, n_local_heads=8, intermediate_size=28672),
    "Mistral-7B": dict(n_layer=32, n_head=32, n_local_heads=8, dim=4096, intermediate_size=14336, vocab_size=32000),
    "stories15M": dict(n_layer=6, n_head=6, dim=288),
    "stories110M": dict(n_layer=12, n_head=12, dim=768),
    "Meta-Llama-3-8B-Instruct": _llama3(8192),
    "Meta-Llama-3.1-8B-Instruct": _llama3(131072, rope_scaling={
        "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
        "original_max_position_embeddings": 8192, "rope_type": "llama3"}),
    "Qwen2-1.5B-Instruct": _qwen2(28, 12, 2, 1536, 8960, 151936),
    "Qwen2-0.5B-Instruct": _qwen2(24, 14, 2, 896, 4864, 151936),
    "Qwen2-7B-Instruct": _qwen2(28, 28, 4, 3584, 18944, 152064),
}


class Transformer(nn.Module):
    def __init__(self, config: ModelArgs) -> None:
        super().__init__()
        self.config = config
        self.tok_embeddings = nn.Embedding(config.vocab_size, config.dim)
        self.layers = nn.ModuleList(TransformerBlock(config) for _ in range(config.n_layer))
        self.norm = RMSNorm(config.dim, eps=config.norm_eps)
        self.output = nn.Linear(config.dim, config.vocab_size, bias=False)
        self.freqs_cis: Optional[Tensor] = None
        self.max_batch_size = 1
        # decode through the fused one-token kernels (5 launches per layer); False = plain tensor ops
        # around the cache kernel (used by the A/B tests)
        self.fused_decode = True
        self._bufs = None
        self.tp_link = None  # tp.enable_fused_exchange: all-reduces fused into the projection kernels
        # decode a token as ONE persistent kernel (csrc/decode_token.cu) when the cache policies allow it;
        # False = the 5-launches-per-layer path above (also the fallback for hybrid / quantised / debug caches)
        self.token_kernel = os.environ.get("CC_TOKEN_KERNEL", "1") != "0"
        self._tplan = None
        self._tplan_why = None

    def setup_caches(self, **kwargs):
        """Per layer: cache + prompt compressor from the strategy names (reference model.py:191-233)."""
        kwargs = dict(kwargs)
        strategies = kwargs.pop("cache_strategy")
        head_dim = self.config.dim // self.config.n_head
        dtype = self.output.weight.dtype
        per_layer = {"max_cache_length", "recent_window", "prompt_compression_strategy"}
        for i, block in enumerate(self.layers):
            ctor, relevant = get_cache_constructor(cache_strategy=strategies[i])
            layer_kwargs = {k: (kwargs[k][i] if k in per_layer else kwargs[k]) for k in relevant}
            att = block.attention
            att.kv_cache = ctor(self.max_batch_size, self.config.n_local_heads, head_dim, dtype, **layer_kwargs)
            att.prompt_compressor = get_prompt_compressor_constructor(kwargs["prompt_compression_strategy"][i])(
                head_specific=att.kv_cache.head_specific, **layer_kwargs)
        self.freqs_cis = precompute_freqs_cis(self.config.block_size, head_dim, self.config.rope_base, dtype,
                                              self.config.rope_scaling)

    def reset_caches(self):
        for layer in self.layers:
            layer.attention.kv_cache.reset()

    def prompt_cache_overflow(self, prompt_length: int):
        return [prompt_length > layer.attention.kv_cache.max_cache_length for layer in self.layers]

    def get_cache_stats(self, prompt_len, gen_len):
        stats, per_key, mem = {}, defaultdict(list), 0.0
        seq_len = torch.tensor(prompt_len + gen_len)
        for i, layer in enumerate(self.layers):
            s = layer.attention.kv_cache.compute_statistics(seq_len=seq_len)
            mem += s.pop("cache_memory_gb")
            for k, v in s.items():
                stats[f"{k}_{i}"] = v
                per_key[k].append(v)
        for k, v in per_key.items():
            stats[f"{k}_avg"] = sum(v) / len(v)
        stats["cache_memory_gb"] = mem
        return stats

    def min_cache_length(self):
        return min(layer.attention.kv_cache.max_cache_length for layer in self.layers)

    # ---- one-token decode through the fused kernels ------------------------------------------
    def _fused_ready(self, idx) -> bool:
        w = self.output.weight
        return (self.fused_decode and idx.is_cuda and w.dtype == torch.bfloat16 and idx.numel() == 1
                and self.output.in_features % 256 == 0
                and all(l.feed_forward.w2.in_features % 256 == 0 and l.attention.wo.in_features % 256 == 0
                        for l in self.layers[:1]))

    def _decode_buffers(self, device):
        if self._bufs is None or self._bufs["x"].device != device:
            att, ffn = self.layers[0].attention, self.layers[0].feed_forward
            mk = lambda n: torch.empty(n, dtype=torch.bfloat16, device=device)  # noqa: E731
            dim = self.output.in_features
            self._bufs = {"x": mk(dim), "h": mk(dim), "part": mk(dim), "q": mk(att.n_head * att.head_dim),
                          "k": mk(att.n_local_heads * att.head_dim), "v": mk(att.n_local_heads * att.head_dim),
                          "gate": mk(ffn.w1.out_features), "logits": mk(self.output.out_features)}
        return self._bufs

    # ---- one token = one kernel ----------------------------------------------------------------
    def token_plan(self):
        """The whole-token kernel's plan for the current weights + caches, or None when some cache needs the
        per-layer kernel (the reason is kept in `_tplan_why`).  Under tensor parallelism building the plan is a
        collective (symmetric-memory rendezvous): every rank gets here in the same decode call."""
        if not (self.token_kernel and self.fused_decode):
            return None
        from .token_kernel import TokenPlan, unsupported_reason
        sig = TokenPlan.signature_of(self)
        if self._tplan is not None and self._tplan.signature == sig:
            return self._tplan
        if self._tplan is None and self._tplan_why == sig:
            return None
        self._tplan = None
        why = unsupported_reason(self)
        if why is None and self.layers[0].attention.tp_group is not None and os.environ.get("CC_TP_NCCL") == "1":
            why = "CC_TP_NCCL=1 selects NCCL all-reduces between the per-layer kernels"
        if why is not None:
            self._tplan_why = sig
            self.token_kernel_reason = why
            return None
        self._tplan = TokenPlan(self)
        return self._tplan

    def decode_token_fused(self, idx: Tensor, input_pos: Tensor):
        """(token int32 [1], probs bf16 [vocab]) of one greedy decode step through the whole-token kernel, or None
        when that kernel cannot serve this model (reference generation_utils.py:163-178 + 136-142)."""
        if idx.numel() != 1 or not idx.is_cuda:
            return None
        plan = self.token_plan()
        if plan is None:
            return None
        tok = idx.view(-1) if idx.dtype == torch.int32 else idx.view(-1).int()
        ip = input_pos if input_pos.dtype == torch.int32 else input_pos.int()
        token, probs, _ = plan.decode(tok, ip)
        return token, probs

    def _decode_fused(self, idx: Tensor, input_pos: Tensor) -> Tensor:
        """embedding -> per layer [norm+wqkv+RoPE | cache step | wo+residual | norm+gate | w2+residual] -> head.
        Same bf16 rounding points as the tensor-op path; every kernel takes `input_pos` from device
        memory, so the whole step is capturable in one CUDA graph."""
        ip = input_pos if input_pos.dtype == torch.int32 else input_pos.int()
        plan = self.token_plan()
        if plan is not None:
            _, _, logits = plan.decode(idx.view(-1) if idx.dtype == torch.int32 else idx.view(-1).int(), ip)
            return logits.view(1, 1, -1)
        b = self._decode_buffers(idx.device)
        eps = self.norm.eps
        b["x"].copy_(self.tok_embeddings(idx).view(-1))
        x, h = b["x"], b["h"]
        link = self.tp_link
        no_tp = ([], [], None)
        if link is not None:
            return self._decode_fused_tp(idx, ip, b, link)
        for layer in self.layers:
            att, ffn = layer.attention, layer.feed_forward
            ops.rmsnorm_qkv_rope(x, layer.attention_norm.weight, layer.attention_norm.eps, att.wqkv.weight,
                                 att.wqkv.bias, self.freqs_cis, ip, b["q"], b["k"], b["v"], att.n_head,
                                 att.n_local_heads, att.head_dim, *no_tp)
            y = att.kv_cache.decode_step(b["q"].view(1, att.n_head, 1, att.head_dim),
                                         b["k"].view(1, att.n_local_heads, 1, att.head_dim),
                                         b["v"].view(1, att.n_local_heads, 1, att.head_dim), ip, input_ids=idx,
                                         overlap_predecessor=True,  # predecessor = this layer's projection
                                         warm_l2=self._warm_ranges(att, ffn))
            if att.tp_group is None:
                ops.gemv_residual(y.view(-1), None, 0.0, att.wo.weight, x, h, *no_tp)
            else:  # row-parallel wo: partial sums, one all-reduce, then the residual (tp.py:156-160)
                ops.gemv_residual(y.view(-1), None, 0.0, att.wo.weight, None, b["part"], *no_tp)
                torch.distributed.all_reduce(b["part"], group=att.tp_group)
                torch.add(x, b["part"], out=h)
            ops.rmsnorm_ffn_gate(h, layer.ffn_norm.weight, layer.ffn_norm.eps, ffn.w1.weight, ffn.w3.weight,
                                 b["gate"], *no_tp)
            if ffn.tp_group is None:
                ops.gemv_residual(b["gate"], None, 0.0, ffn.w2.weight, h, x, *no_tp)
            else:
                ops.gemv_residual(b["gate"], None, 0.0, ffn.w2.weight, None, b["part"], *no_tp)
                torch.distributed.all_reduce(b["part"], group=ffn.tp_group)
                torch.add(h, b["part"], out=x)
        ops.gemv_residual(x, self.norm.weight, eps, self.output.weight, None, b["logits"], *no_tp)
        return b["logits"].view(1, 1, -1)

    def _decode_fused_tp(self, idx, ip, b, link):
        """Tensor-parallel decode with the all-reduces fused into the projections: the row-parallel wo / w2
        kernels write their partial outputs into every rank's exchange buffer over NVLink (op 2l / 2l+1),
        and the next kernel sums them (+ the residual) in its prologue.  No NCCL call, no extra kernel."""
        x, h = b["x"], b["h"]
        lt = link.tensors()
        prev = -1  # exchange op whose result is this layer's input (the previous layer's w2)
        for l, layer in enumerate(self.layers):
            att, ffn = layer.attention, layer.feed_forward
            # input: x (layer 0) or h_prev + allreduce(w2 partials) -> written to x by CTA 0
            ops.rmsnorm_qkv_rope(x if prev < 0 else h, layer.attention_norm.weight, layer.attention_norm.eps,
                                 att.wqkv.weight, att.wqkv.bias, self.freqs_cis, ip, b["q"], b["k"], b["v"],
                                 att.n_head, att.n_local_heads, att.head_dim,
                                 lt if prev >= 0 else [], link.meta(recv_op=prev), x if prev >= 0 else None)
            y = att.kv_cache.decode_step(b["q"].view(1, att.n_head, 1, att.head_dim),
                                         b["k"].view(1, att.n_local_heads, 1, att.head_dim),
                                         b["v"].view(1, att.n_local_heads, 1, att.head_dim), ip, input_ids=idx,
                                         overlap_predecessor=True)
            ops.gemv_residual(y.view(-1), None, 0.0, att.wo.weight, None, b["part"], lt, link.meta(send_op=2 * l), None)
            # h = x + allreduce(wo partials), then ffn_norm -> gate
            ops.rmsnorm_ffn_gate(x, layer.ffn_norm.weight, layer.ffn_norm.eps, ffn.w1.weight, ffn.w3.weight, b["gate"],
                                 lt, link.meta(recv_op=2 * l), h)
            ops.gemv_residual(b["gate"], None, 0.0, ffn.w2.weight, None, b["part"], lt, link.meta(send_op=2 * l + 1), None)
            prev = 2 * l + 1
        ops.gemv_residual(h, self.norm.weight, self.norm.eps, self.output.weight, None, b["logits"],
                          lt, link.meta(recv_op=prev), x)
        link.epoch.add_(1)
        return b["logits"].view(1, 1, -1)

    # Optional L2 warm-up: bytes of w1 / w3 rows (plus all of wo) the cache step asks L2 to pre-load while it
    # waits on its own latencies (cc_decode_step prefetch hints).  Measured on B200 (profiles/r1_notes.md):
    # 333 tok/s without, 332 / 331 / 329 / 325 tok/s with wo / +10 / +24 / +40 MB -- no gain, so it is off.
    warm_l2_ffn_bytes = -1

    def _warm_ranges(self, att, ffn):
        if self.warm_l2_ffn_bytes < 0:
            return ()
        rows = max(0, min(ffn.w1.out_features, self.warm_l2_ffn_bytes // (ffn.w1.in_features * 2)))
        out = [att.wo.weight]
        if rows:
            out += [ffn.w1.weight[:rows], ffn.w3.weight[:rows]]
        return out

    def kernels_per_token(self) -> int:
        """Launches of this repo's kernels in one fused decode step (bench.py: gpu_launches)."""
        if self._tplan is not None:
            return 1
        return 5 * len(self.layers) + 1

    def forward(self, idx: Tensor, input_pos: Tensor, is_prefill, mask: Optional[Tensor] = None,
                attn_top_k: Optional[float] = 1.0) -> Tensor:
        assert self.freqs_cis is not None, "Caches must be initialized first"
        if not is_prefill and self._fused_ready(idx):
            # the fused kernels attend over every live slot: top-k decode attention is refused, not ignored
            assert attn_top_k in (None, 1, 1.0), "top-k decode attention is not supported with cache masks"
            return self._decode_fused(idx, input_pos)
        freqs_cis = self.freqs_cis[input_pos]
        x = self.tok_embeddings(idx)
        for layer in self.layers:
            x = layer(x, idx, input_pos, is_prefill, freqs_cis, mask, attn_top_k=attn_top_k)
        if is_prefill:
            x = x[:, -1:]  # only the last position feeds the sampler (generation_utils.greedy reads logits[0, -1])
        return self.output(self.norm(x))

    @classmethod
    def from_name(cls, name: str):
        return cls(ModelArgs.from_name(name))


class TransformerBlock(nn.Module):
    def __init__(self, config: ModelArgs) -> None:
        super().__init__()
        self.attention = Attention(config)
        self.feed_forward = FeedForward(config)
        self.ffn_norm = RMSNorm(config.dim, config.norm_eps)
        self.attention_norm = RMSNorm(config.dim, config.norm_eps)

    def forward(self, x, input_ids, input_pos, is_prefill, freqs_cis, mask, attn_top_k=1.0):
        h = x + self.attention(self.attention_norm(x), input_ids, freqs_cis, mask, is_prefill, input_pos,
                               attn_top_k=attn_top_k)
        return h + self.feed_forward(self.ffn_norm(h))


class Attention(nn.Module):
    def __init__(self, config: ModelArgs):
        super().__init__()
        assert config.dim % config.n_head == 0
        qkv_out = (config.n_head + 2 * config.n_local_heads) * config.head_dim
        self.wqkv = nn.Linear(config.dim, qkv_out, bias=config.attention_bias)
        self.wo = nn.Linear(config.dim, config.dim, bias=False)
        self.kv_cache = None
        self.prompt_compressor = None
        self.n_head, self.head_dim = config.n_head, config.head_dim
        self.n_local_heads, self.dim = config.n_local_heads, config.dim
        self.tp_group = None  # set by tp.apply_tp: all-reduce of the row-parallel wo output
        self._register_load_state_dict_pre_hook(self.load_hook)

    def load_hook(self, state_dict, prefix, *args):
        if prefix + "wq.weight" in state_dict:  # HF-style separate projections -> fused wqkv
            parts = [state_dict.pop(prefix + f"w{n}.weight") for n in "qkv"]
            state_dict[prefix + "wqkv.weight"] = torch.cat(parts)

    def compress_prompt(self, input_pos, k_val, v_val, attn):
        if self.kv_cache.max_cache_length < input_pos.shape[0]:
            return self.prompt_compressor(input_pos, k_val, v_val, attn=attn)
        return input_pos, k_val, v_val, attn

    def _project(self, x, freqs_cis):
        bsz, seqlen, _ = x.shape
        kv = self.n_local_heads * self.head_dim
        q, k, v = self.wqkv(x).split([self.dim, kv, kv], dim=-1)
        q = apply_rotary_emb(q.view(bsz, seqlen, self.n_head, self.head_dim), freqs_cis).transpose(1, 2)
        k = apply_rotary_emb(k.view(bsz, seqlen, self.n_local_heads, self.head_dim), freqs_cis).transpose(1, 2)
        v = v.view(bsz, seqlen, self.n_local_heads, self.head_dim).transpose(1, 2)
        return q, k, v

    def forward(self, x: Tensor, input_ids: Tensor, freqs_cis: Tensor, mask: Tensor, is_prefill: bool,
                input_pos: Optional[Tensor] = None, attn_top_k: Optional[float] = 1.0) -> Tensor:
        bsz, seqlen, _ = x.shape
        q, k, v = self._project(x, freqs_cis)
        if not is_prefill:
            assert attn_top_k in (None, 1.0), "top-k decode attention is not supported with cache masks"
            # eviction + overwrite + attention + pooling + score update: one kernel
            y = self.kv_cache.decode_step(q.contiguous(), k.contiguous(), v.contiguous(), input_pos,
                                          input_ids=input_ids)
        else:
            y = self._prefill(q, k, v, input_ids, input_pos, mask)
        y = self.wo(y.transpose(1, 2).reshape(bsz, seqlen, self.dim))
        if self.tp_group is not None:
            torch.distributed.all_reduce(y, group=self.tp_group)
        return y

    def _prefill(self, q, k, v, input_ids, input_pos, mask):
        """Causal attention over the prompt, prompt compression, cache fill, score seeding
        (reference model.py:399-427, prefill branch)."""
        cache = self.kv_cache
        kwargs = {"input_ids": input_ids}
        if cache.return_attn():
            windows = cache.profiling_windows(q.shape[2]) if hasattr(cache, "profiling_windows") else ()
            y, attn = prefill_attention_stats(q, k, v, windows=windows)
        else:
            y, attn = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True), None
        input_pos, k, v, attn = self.compress_prompt(input_pos, k, v, attn)
        cache.update_kv(input_pos, k, v, True, **kwargs)
        if isinstance(attn, PrefillAttnStats) and not hasattr(cache, "profile_and_update") and not hasattr(cache, "compressed"):
            # prompt fit the cache: seed with the column means directly
            seq_len = attn.col_sum.shape[-1]
            attn = attn.col_sum / (seq_len - input_pos)
        cache.update_state(input_pos, k, v, True, attn, **kwargs)
        return y


class FeedForward(nn.Module):
    def __init__(self, config: ModelArgs) -> None:
        super().__init__()
        self.w1 = nn.Linear(config.dim, config.intermediate_size, bias=False)
        self.w3 = nn.Linear(config.dim, config.intermediate_size, bias=False)
        self.w2 = nn.Linear(config.intermediate_size, config.dim, bias=False)
        self.tp_group = None

    def forward(self, x: Tensor) -> Tensor:
        y = self.w2(F.silu(self.w1(x)) * self.w3(x))
        if self.tp_group is not None:
            torch.distributed.all_reduce(y, group=self.tp_group)
        return y


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x: Tensor) -> Tensor:
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x) * self.weight


def precompute_freqs_cis(seq_len: int, n_elem: int, base: int = 10000, dtype: torch.dtype = torch.bfloat16,
                         rope_scaling: Optional[Dict[str, Any]] = None) -> Tensor:
    """[seq_len, n_elem/2, 2] (cos, sin) table, optional Llama-3.1 frequency scaling (reference model.py:460-504)."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: n_elem // 2].float() / n_elem))
    if rope_scaling is not None:
        assert rope_scaling["rope_type"] == "llama3", "Only Llama 3.1 scaling is supported"
        orig = rope_scaling["original_max_position_embeddings"]
        lo, hi, factor = rope_scaling["low_freq_factor"], rope_scaling["high_freq_factor"], rope_scaling["factor"]
        wavelen = 2 * math.pi / freqs
        smooth = ((orig / wavelen - lo) / (hi - lo)).clamp(0, 1)
        blended = (1 - smooth) * freqs / factor + smooth * freqs
        freqs = torch.where(wavelen < orig / hi, freqs, torch.where(wavelen > orig / lo, freqs / factor, blended))
    angles = torch.outer(torch.arange(seq_len, device=freqs.device).float(), freqs)
    return torch.stack([torch.cos(angles), torch.sin(angles)], dim=-1).to(dtype=dtype)


def apply_rotary_emb(x: Tensor, freqs_cis: Tensor) -> Tensor:
    """Rotate adjacent pairs in fp32, cast back (reference model.py:507-519)."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    fc = freqs_cis.view(1, xs.size(1), 1, xs.size(3), 2)
    re = xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1]
    im = xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]
    return torch.stack([re, im], -1).flatten(3).type_as(x)
