#!/usr/bin/env python
"""Whole-model trace of BASELINE config 1, recorded by EXECUTING THE UNMODIFIED REFERENCE on CPU (build container
only, needs /root/reference and ~40 GB of RAM):

    python tests/golden/make_golden_model.py            # writes tests/golden/model_cfg1.npz

Llama-3-8B shape (32 layers, bf16, weights from tests/model_weights.py), recent_global cache with 64 slots per layer,
128-token prompt, 16 generated tokens through the reference's own generate() (generation_utils.py:399-531:
setup_caches -> prefill -> decode_n_tokens -> decode_one_token, greedy).  Recorded: the prompt, the generated tokens,
per decode step and layer the slot `_fill` wrote (the eviction trace), every layer's `pos` after the run, and per step
the 64 largest probabilities with their token ids (a random-init model over a 128k vocabulary is nearly flat at the top:
bf16 probabilities tie, so tokens are compared through these margins, not by equality).
"""
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from model_weights import seeded_weights  # noqa: E402

MODEL, L, P, T = "Meta-Llama-3-8B-Instruct", 64, 128, 16


def cache_kwargs():
    return dict(max_cache_length=[float(L)], cache_bits=None, cache_length_pattern="tile", cache_strategy=["recent_global"],
                cache_strategy_pattern="tile", feed_long_prompts=False, prompt_compression_strategy=["recent_global"],
                global_tokens=4, recent_window=10, history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)


def main():
    sys.path.insert(0, str(REF))
    import generation_utils as rgu
    import model as rmodel
    torch.set_num_threads(8)
    torch.cuda.max_memory_reserved = lambda *a, **k: 0   # generate() reports it; there is no GPU here
    with torch.device("meta"):
        m = rmodel.Transformer.from_name(MODEL)
    m = m.to(dtype=torch.bfloat16).to_empty(device="cpu")
    seeded_weights(m)
    m.eval()
    rgu.setup_caches(m, None, torch.device("cpu"), P + T, cache_kwargs())
    trace = []   # (layer, slot) in call order

    def spy(layer, cache):
        real = cache._fill

        def fill(input_pos, k_val, v_val, fill_idxs, **kw):
            if torch.is_tensor(fill_idxs) and fill_idxs.numel() == 1:
                trace.append((layer, int(fill_idxs)))
            return real(input_pos, k_val, v_val, fill_idxs, **kw)
        cache._fill = fill
    for i, b in enumerate(m.layers):
        spy(i, b.attention.kv_cache)
    prompt = torch.randint(0, m.config.vocab_size, (P,), dtype=torch.int32, generator=torch.Generator().manual_seed(2024))
    prefill, decode_one_token = rgu.compile_funcs(False)
    with torch.no_grad():
        seq, probs, stats = rgu.generate(m, prompt, prefill, decode_one_token, max_new_tokens=T)
    n_layer = len(m.layers)
    steps = len(trace) // n_layer
    assert steps * n_layer == len(trace) and steps == T - 1, (len(trace), steps)
    fill = np.zeros((steps, n_layer), dtype=np.int32)
    for i, (layer, slot) in enumerate(trace):
        assert layer == i % n_layer
        fill[i // n_layer, layer] = slot
    top_p = np.zeros((T, 64), dtype=np.float32)
    top_i = np.zeros((T, 64), dtype=np.int32)
    for t, pr in enumerate(probs):
        v, ix = torch.topk(pr.float().view(-1), 64)
        top_p[t], top_i[t] = v.numpy(), ix.numpy()
    pos = np.stack([b.attention.kv_cache.pos[0, 0].numpy() for b in m.layers]).astype(np.int32)
    np.savez_compressed(HERE / "model_cfg1.npz", prompt=prompt.numpy(), seq=seq.numpy().astype(np.int32), fill_idx=fill,
                        final_pos=pos, top_p=top_p, top_i=top_i)
    print("generated", seq[P:].tolist(), "decode tok/s", stats["decode_toks_per_sec"])
    print("top-2 probability ratio per step", (top_p[:, 1] / top_p[:, 0]).round(3).tolist())


if __name__ == "__main__":
    main()
