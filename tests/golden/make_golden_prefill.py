#!/usr/bin/env python
"""Generate golden PREFILL vectors by EXECUTING THE UNMODIFIED REFERENCE on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_prefill.py      # writes tests/golden/prefill_*.npz

A reference `model.Attention` (model.py:330-432) with a reference cache and prompt compressor is
driven through ONE prefill on seeded inputs.  Recorded at the boundary of the prefill attention:

  q, k, v        what scaled_dot_product_attention receives (post-RoPE, k/v before the GQA expand)
  y              its output
  col_sum        pooled.sum(2)                         (prompt_compression.py:192, cache.py:704)  bf16
  obs_mean       pooled[:, :, -16:, :].mean(2)          (prompt_compression.py:175)               bf16
  rows_idx/rows  a few full rows of the pooled map (model.py:413-418) for direct probability checks
  priority, keep_idxs, cum_attn   the SnapKV compressor's outputs for the recorded prompt
                                  (prompt_compression.py:148-194), when the cache returns attention

bf16 tensors are stored as their uint16 bit patterns.  Nothing from the reference is copied into
the repo; these files plus this script are the pin for oracle/prefill.py and the CUDA prefill kernel.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

FIXTURES = {
    # name: strategy decides return_attn (heavy_hitter: naive path + statistics, l2: F.sdpa MATH path)
    "prefill_hh_g4_d128": dict(strategy="heavy_hitter", Hkv=2, G=4, D=128, S=320, L=128, x_scale=2.0),
    "prefill_hh_g7_ragged": dict(strategy="heavy_hitter", Hkv=2, G=7, D=128, S=203, L=96, x_scale=2.5, bias=True),
    "prefill_hh_g2_d64": dict(strategy="heavy_hitter", Hkv=2, G=2, D=64, S=150, L=64, x_scale=2.0),
    "prefill_hh_short": dict(strategy="heavy_hitter", Hkv=1, G=4, D=128, S=9, L=64, x_scale=2.0),
    "prefill_l2_math": dict(strategy="l2", Hkv=2, G=4, D=128, S=260, L=96, x_scale=2.0),
}


def u16(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().view(torch.uint16).numpy().copy()


def make(name: str, spec: dict, seed: int = 4321) -> dict:
    sys.path.insert(0, str(REF))
    import model as ref_model  # noqa: the reference, unmodified
    from cache import get_cache_constructor
    from prompt_compression import get_prompt_compressor_constructor

    torch.manual_seed(seed)
    strategy, Hkv, G, D, S, L = (spec[k] for k in ("strategy", "Hkv", "G", "D", "S", "L"))
    n_head, dim = Hkv * G, Hkv * G * D
    cfg = ref_model.ModelArgs(block_size=S + 8, vocab_size=1000, n_layer=1, n_head=n_head, dim=dim,
                              n_local_heads=Hkv, intermediate_size=256, attention_bias=spec.get("bias", False),
                              rope_base=500000)
    attn = ref_model.Attention(cfg)
    with torch.no_grad():
        for p in attn.parameters():
            p.normal_(0.0, 0.02)
    attn = attn.to(torch.bfloat16).eval()
    kwargs = dict(max_cache_length=L, global_tokens=4, max_seq_length=S + 8, cache_bits=None, recent_window=10,
                  history_window_size=1, attn_thresholding=False, prompt_compression_strategy=strategy)
    cls, relevant = get_cache_constructor(strategy)
    cache = cls(1, Hkv, D, torch.bfloat16, **{k: kwargs[k] for k in relevant})
    attn.kv_cache = cache
    attn.prompt_compressor = get_prompt_compressor_constructor(strategy)(
        head_specific=cache.head_specific, **{k: kwargs[k] for k in relevant})
    freqs = ref_model.precompute_freqs_cis(cfg.block_size, D, cfg.rope_base, torch.bfloat16)

    out: dict = {}
    real_sdpa = ref_model.scaled_dot_product_attention

    def sdpa_spy(query, key, value, **kw):
        y, p = real_sdpa(query, key, value, **kw)
        # key/value arrive GQA-expanded (model.py:395-400): keep one copy per KV head
        out["q__bf16"], out["k__bf16"], out["v__bf16"] = u16(query[0]), u16(key[0, ::G]), u16(value[0, ::G])
        out["y__bf16"] = u16(y[0])
        return y, p

    real_compress = attn.compress_prompt

    def compress_spy(input_pos, k_val, v_val, pooled):
        if pooled is not None:
            out["col_sum__bf16"] = u16(pooled.sum(dim=2)[0])
            obs = min(16, S)
            out["obs_mean__bf16"] = u16(pooled[:, :, -obs:, :].mean(dim=2)[0])
            rows = sorted({0, 1, S // 3, S // 2, S - 17 if S > 17 else 0, S - 2, S - 1} & set(range(S)))
            out["rows_idx"] = np.asarray(rows, dtype=np.int32)
            out["rows__bf16"] = u16(pooled[0][:, rows, :])
        res = real_compress(input_pos, k_val, v_val, pooled)
        if pooled is not None and S > L:
            keep_idxs, _, _, cum = res
            out["keep_idxs"] = keep_idxs.numpy().astype(np.int32)
            out["cum_attn__bf16"] = u16(cum[0])
            pr = attn.prompt_compressor._token_importances(input_pos, k_val, v_val, attn=pooled)
            out["priority__bf16"] = u16(pr[0])
        return res

    ref_model.scaled_dot_product_attention = sdpa_spy
    attn.compress_prompt = compress_spy
    with torch.no_grad():
        ids = torch.randint(0, 1000, (1, S), dtype=torch.int32)
        x = (torch.randn(1, S, dim) * spec["x_scale"]).to(torch.bfloat16)
        input_pos = torch.arange(S)
        causal = torch.tril(torch.ones(S, S, dtype=torch.bool)).view(1, 1, S, S)
        attn(x, ids, freqs[input_pos], causal, True, input_pos)
    ref_model.scaled_dot_product_attention = real_sdpa
    meta = dict(spec, name=name, seed=seed, torch=torch.__version__, return_attn=bool(cache.return_attn()),
                global_tokens=4, obs_len=min(16, S))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def main():
    assert REF.exists(), "needs the reference checkout at /root/reference (build container only)"
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(FIXTURES)):
        data = make(name, FIXTURES[name])
        path = OUT / f"{name}.npz"
        np.savez_compressed(path, **data)
        print(f"{name}: {path.stat().st_size / 1024:.0f} KiB  keys={sorted(k for k in data if k != 'meta')}")


if __name__ == "__main__":
    main()
