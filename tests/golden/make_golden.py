#!/usr/bin/env python
"""Generate golden decode traces by EXECUTING THE UNMODIFIED REFERENCE on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

For each fixture a reference `model.Attention` module (model.py:330-432) with a reference
KV cache (cache.py) and prompt compressor is driven through one prefill and T decode steps
on seeded random inputs.  What the hot path sees and produces is recorded at its boundary:

  per step : q (post-RoPE, sdpa `query`), k_new / v_new (update_kv args), input_pos, input_id,
             fill_idx (`_fill` fill_idxs = the eviction trace), y (sdpa output),
             probs (what update_state receives, GQA-pooled), pos after the step
  once     : every cache buffer after prefill (`init_*`) and after the last step (`final_*`)

bf16 tensors are stored as their uint16 bit patterns.  Nothing from the reference is copied
into the repo; the committed .npz files plus this script are the pin for oracle/cc_oracle.c.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

HYBRID_YAML = [  # cache_configs/hybrid.yaml
    {"strategy": "window", "recent_window": 0.1},
    {"strategy": "window_heavy_hitter", "heavy_hitter_frac": 0.25, "recent_window": 0.1},
    {"strategy": "window_heavy_hitter", "heavy_hitter_frac": 0.5, "recent_window": 0.1},
    {"strategy": "full"},
]
FASTGEN_YAML = [  # cache_configs/fastgen.yaml
    {"strategy": "special"},
    {"strategy": "special_punc"},
    {"strategy": "special_punc_heavy_hitter", "heavy_hitter_frac": 0.3},
    {"strategy": "special_punc_heavy_hitter_window", "recent_window": 0.3, "heavy_hitter_frac": 0.3},
    {"strategy": "full"},
]

FIXTURES = {
    # name: dict(strategy, Hkv, G, D, L, S(prompt), T(decode steps), extra kwargs)
    "hh_fill_then_evict": dict(strategy="heavy_hitter", Hkv=2, G=4, D=128, L=64, S=48, T=150),
    "hh_snapkv": dict(strategy="heavy_hitter", Hkv=4, G=4, D=128, L=128, S=300, T=100),
    "hh_g7_bias": dict(strategy="heavy_hitter", Hkv=2, G=7, D=128, L=96, S=40, T=120, bias=True),
    "hh_d64_ragged": dict(strategy="heavy_hitter", Hkv=2, G=2, D=64, L=40, S=25, T=100),
    # heavy hitter with a bounded history window (bf16 ring, cache.py:661-667) and with thresholded (bool) history
    "hh_window4": dict(strategy="heavy_hitter", Hkv=2, G=4, D=128, L=64, S=100, T=120, history_window_size=4),
    # attn_thresholding=True cannot be recorded: the unmodified reference raises in update_state on this image's
    # torch 2.11 ("Index put requires the source and destination dtypes match, got Bool ... Int", cache.py:708-721),
    # so the thresholded variant is pinned on the oracle's reading of the code only (tests/test_gpu_parity.py).
    # debug_<strategy> (KVCacheAnalysis, cache.py:1291-1420): full cache + shadow heavy-hitter cache + attention loss
    "debug_hh": dict(strategy="debug_heavy_hitter", Hkv=2, G=4, D=128, L=32, S=60, T=50),
    # quantised KV (--cache_bits): the whole cache is de-quantised, updated and re-quantised every step
    # (cache.py:283-338, quantization_utils.py), groups = one slot across all heads
    "hh_bits8": dict(strategy="heavy_hitter", Hkv=2, G=4, D=128, L=48, S=80, T=80, cache_bits=8),
    "hh_bits4": dict(strategy="heavy_hitter", Hkv=2, G=2, D=64, L=40, S=30, T=80, cache_bits=4),
    "rg_bits2": dict(strategy="recent_global", Hkv=2, G=4, D=128, L=32, S=60, T=60, cache_bits=2),
    # BASELINE.json per-layer cache shapes (Llama-3-8B heads): config 1 exactly; configs 2 / 3 with a shorter prompt
    # (the cache sees the same 1024 / 512 slots; the reference's CPU prefill materialises the S x S map)
    "cfg1_rg_l64": dict(strategy="recent_global", Hkv=8, G=4, D=128, L=64, S=128, T=16),
    "cfg2_hh_l1024": dict(strategy="heavy_hitter", Hkv=8, G=4, D=128, L=1024, S=1536, T=24),
    "cfg3_l2_l512": dict(strategy="l2", Hkv=8, G=4, D=128, L=512, S=768, T=24),
    "rg_compress": dict(strategy="recent_global", Hkv=2, G=4, D=128, L=64, S=128, T=100),
    "l2_compress": dict(strategy="l2", Hkv=4, G=2, D=128, L=64, S=100, T=120),
    "l2_fill": dict(strategy="l2", Hkv=2, G=4, D=64, L=48, S=20, T=80),
    "full_external": dict(strategy="full", Hkv=2, G=4, D=128, L=64, S=20, T=30),
    # per-head q scaling makes heads differently peaky so the reference's profiling
    # (cache.py:1138-1187) picks a MIX of strategies
    "hybrid_window_hh": dict(strategy="hybrid", Hkv=4, G=7, D=128, L=240, S=140, T=100,
                             hybrid_strategies=HYBRID_YAML, min_recovery_frac=0.55, x_scale=2.0,
                             head_q_scales=[0.25, 1.0, 2.0, 4.0]),
    "hybrid_fastgen": dict(strategy="hybrid", Hkv=4, G=2, D=128, L=224, S=128, T=96,
                           hybrid_strategies=FASTGEN_YAML, min_recovery_frac=0.5, x_scale=2.0,
                           head_q_scales=[0.25, 1.0, 2.0, 4.0],
                           vocab=40, special=[[0], [1]], punctuation=[5, 6, 7]),
    "hybrid_window_forced": dict(strategy="hybrid", Hkv=4, G=7, D=128, L=240, S=140, T=100,
                                 hybrid_strategies=HYBRID_YAML, min_recovery_frac=0.55, x_scale=2.0,
                                 force_strategies=[0, 1, 2, 3]),
    # strategies overridden after the reference's own prefill so that the special-only /
    # special_punc heads (token-dropped path, cache.py:948-950) are exercised at decode
    "hybrid_fastgen_forced": dict(strategy="hybrid", Hkv=4, G=2, D=128, L=224, S=128, T=96,
                                  hybrid_strategies=FASTGEN_YAML, min_recovery_frac=0.5, x_scale=2.0,
                                  force_strategies=[0, 1, 3, 4],
                                  vocab=40, special=[[0], [1]], punctuation=[5, 6, 7]),
}


def u16(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.uint16).numpy().copy()


def snapshot(cache, prefix: str, out: dict):
    for name, buf in cache.named_buffers():
        if name in ("special_ids", "punc_ids"):
            continue
        t = buf.detach()
        key = f"{prefix}_{name}"
        if t.dtype == torch.bfloat16:
            out[key + "__bf16"] = u16(t)
        elif t.dtype == torch.bool:
            out[key] = t.numpy().astype(np.uint8)
        else:
            out[key] = t.numpy().copy()
    if getattr(cache, "quantize", False):
        from quantization_utils import dequantize_tensor
        for n in ("k", "v"):
            dq = dequantize_tensor(getattr(cache, f"{n}_cache"), getattr(cache, f"{n}_scales"),
                                   getattr(cache, f"{n}_zero_points"), cache.cache_shape, n_bit=cache.n_bit, axis=2)
            out[f"{prefix}_{n}_cache_dq__bf16"] = u16(dq)
            out[f"{prefix}_{n}_scales__bf16"] = u16(getattr(cache, f"{n}_scales"))
            out[f"{prefix}_{n}_zero_points__bf16"] = u16(getattr(cache, f"{n}_zero_points"))
    # Hybrid replaces these buffers by plain attributes during profiling (cache.py:1147,1151,1240)
    for name in ("cache_cts", "num_special", "num_punc"):
        if hasattr(cache, name) and torch.is_tensor(getattr(cache, name)):
            out[f"{prefix}_{name}"] = getattr(cache, name).detach().reshape(-1).to(torch.int32).numpy().copy()


def make(name: str, spec: dict, seed: int = 1234) -> dict:
    sys.path.insert(0, str(REF))
    import model as ref_model  # noqa: the reference, unmodified
    from cache import get_cache_constructor
    from prompt_compression import get_prompt_compressor_constructor
    from torch.nn.attention import SDPBackend, sdpa_kernel

    torch.manual_seed(seed)
    strategy, Hkv, G, D, L, S, T = (spec[k] for k in ("strategy", "Hkv", "G", "D", "L", "S", "T"))
    n_head, dim = Hkv * G, Hkv * G * D
    cfg = ref_model.ModelArgs(block_size=S + T + 8, vocab_size=spec.get("vocab", 1000), n_layer=1, n_head=n_head,
                              dim=dim, n_local_heads=Hkv, intermediate_size=256,
                              attention_bias=spec.get("bias", False), rope_base=500000)
    attn = ref_model.Attention(cfg)
    with torch.no_grad():
        for p in attn.parameters():
            p.normal_(0.0, 0.02)
        if spec.get("head_q_scales"):
            for h, sc in enumerate(spec["head_q_scales"]):
                attn.wqkv.weight[h * G * D:(h + 1) * G * D] *= sc
    attn = attn.to(torch.bfloat16).eval()

    prompt_strategy = {"heavy_hitter": "heavy_hitter", "l2": "l2", "recent_global": "recent_global",
                       "full": "full", "hybrid": "full", "debug_heavy_hitter": "heavy_hitter"}[strategy]
    kwargs = dict(max_cache_length=L, global_tokens=4, max_seq_length=S + T, cache_bits=spec.get("cache_bits"),
                  recent_window=10, history_window_size=spec.get("history_window_size", 1),
                  attn_thresholding=spec.get("attn_thresholding", False),
                  prompt_compression_strategy=prompt_strategy)
    if strategy == "hybrid":
        kwargs.update(hybrid_strategies=spec["hybrid_strategies"], min_recovery_frac=spec["min_recovery_frac"],
                      token_ids={"special": spec.get("special", [[0], [1]]),
                                 "punctuation": spec.get("punctuation", [5])})
    cls, relevant = get_cache_constructor(strategy)
    if strategy.startswith("debug"):
        # the reference cannot construct KVCacheAnalysis as shipped: its full-cache kwargs omit cache_bits, which
        # KVCache.__init__ reads (cache.py:181, 1317-1324).  A class attribute supplies the missing default; every
        # line of reference code still runs unmodified.
        import cache as ref_cache_mod
        ref_cache_mod.KVCacheAnalysis.cache_bits = None
    cache = cls(1, Hkv, D, torch.bfloat16, **{k: kwargs[k] for k in relevant})
    attn.kv_cache = cache
    attn.prompt_compressor = get_prompt_compressor_constructor(prompt_strategy)(
        head_specific=cache.head_specific, **{k: kwargs[k] for k in relevant})
    freqs = ref_model.precompute_freqs_cis(cfg.block_size, D, cfg.rope_base, torch.bfloat16)

    rec = {k: [] for k in ("q", "k_new", "v_new", "y", "probs", "fill_idx", "pos_after", "input_pos", "input_id",
                           "cpos_after")}
    state = {"decode": False}

    real_sdpa = ref_model.scaled_dot_product_attention

    def sdpa_spy(query, key, value, **kw):
        y, p = real_sdpa(query, key, value, **kw)
        if state["decode"]:
            rec["q"].append(u16(query[0, :, 0, :]))
            rec["y"].append(u16(y[0, :, 0, :]))
        return y, p

    ref_model.scaled_dot_product_attention = sdpa_spy
    real_update_kv, real_fill, real_update_state = cache.update_kv, cache._fill, cache.update_state

    def update_kv_spy(input_pos, k_val, v_val, is_prefill, **kw):
        if not is_prefill:
            rec["k_new"].append(u16(k_val[0, :, 0, :]))
            rec["v_new"].append(u16(v_val[0, :, 0, :]))
        return real_update_kv(input_pos, k_val, v_val, is_prefill, **kw)

    def fill_spy(input_pos, k_val, v_val, fill_idxs, **kw):
        if state["decode"]:
            rec["fill_idx"].append(fill_idxs.detach().reshape(-1).to(torch.int64).numpy().copy())
        return real_fill(input_pos, k_val, v_val, fill_idxs, **kw)

    def update_state_spy(input_pos, k_val, v_val, is_prefill, attn_probs, **kw):
        if state["decode"] and attn_probs is not None:
            rec["probs"].append(u16(attn_probs[0, :, 0, :]))
        return real_update_state(input_pos, k_val, v_val, is_prefill, attn_probs, **kw)

    cache.update_kv, cache._fill, cache.update_state = update_kv_spy, fill_spy, update_state_spy

    xs = spec.get("x_scale", 1.5)
    vocab = cfg.vocab_size
    out: dict = {}
    with torch.no_grad():
        ids = torch.randint(0, vocab, (1, S), dtype=torch.int32)
        x = (torch.randn(1, S, dim) * xs).to(torch.bfloat16)
        input_pos = torch.arange(S)
        causal = torch.tril(torch.ones(S, S, dtype=torch.bool)).view(1, 1, S, S)
        attn(x, ids, freqs[input_pos], causal, True, input_pos)
        if spec.get("force_strategies"):
            cache.cache_strategies = torch.tensor(spec["force_strategies"])
        snapshot(cache, "init", out)
        if strategy == "hybrid":
            out["init_cache_strategies"] = cache.cache_strategies.to(torch.int32).numpy().copy()
            out["requires_heavy_hitter"] = np.array([int(cache.requires_heavy_hitter)])
            print(f"  {name}: strategies {cache.cache_strategies.tolist()} cts {cache.cache_cts.tolist()}")
        state["decode"] = True
        pos_t = torch.tensor([S], dtype=torch.int32)
        for t in range(T):
            tok = torch.randint(0, vocab, (1, 1), dtype=torch.int32)
            x = (torch.randn(1, 1, dim) * xs).to(torch.bfloat16)
            with sdpa_kernel([SDPBackend.MATH]):  # generation_utils.py:194-196
                attn(x, tok, freqs[pos_t], None, False, pos_t)
            rec["input_pos"].append(int(pos_t))
            rec["input_id"].append(int(tok))
            rec["pos_after"].append(cache.pos[0].numpy().astype(np.int16).copy())
            if hasattr(cache, "compressed"):
                rec["cpos_after"].append(cache.compressed.pos[0].numpy().astype(np.int16).copy())
            pos_t += 1
        snapshot(cache, "final", out)
    ref_model.scaled_dot_product_attention = real_sdpa

    for k, v in rec.items():
        if v:
            out[k] = np.stack(v) if isinstance(v[0], np.ndarray) else np.asarray(v, dtype=np.int32)
    meta = dict(spec, name=name, seed=seed, torch=torch.__version__, global_tokens=4, recent_window=10,
                return_attn=bool(cache.return_attn()), head_specific=bool(cache.head_specific))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*", default=list(FIXTURES))
    args = ap.parse_args()
    assert REF.exists(), "needs the reference checkout at /root/reference (build container only)"
    torch.set_num_threads(8)
    for name in args.names:
        data = make(name, FIXTURES[name])
        path = OUT / f"{name}.npz"
        np.savez_compressed(path, **data)
        print(f"{name}: {path.stat().st_size / 1024:.0f} KiB, steps={len(data['input_pos'])}")


if __name__ == "__main__":
    main()
