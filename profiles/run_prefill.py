#!/usr/bin/env python
"""Time the tcgen05 prefill attention kernel (cc_prefill_attention) on the BASELINE head shapes.

    python profiles/run_prefill.py [--once] [--shape llama8k|llama4k|qwen16k]

Reports the kernel time (CUDA events, L2 flushed between iterations) and the achieved tensor throughput on
ALGORITHMIC flops = causal Q.K^T + P.V = 2 * S^2 * D flop per query head (the kernel executes 1.5x that: the
logits are computed twice, once for the softmax statistics and once for the probabilities).  For context:
torch's fused F.scaled_dot_product_attention (no probabilities, no statistics) and the torch restatement.
--once: a single launch per shape (for ncu)."""
import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cold_compress_b200.attention_utils import prefill_attention_stats  # noqa: E402

SHAPES = {"llama8k": (8, 4, 128, 7680, ()), "llama4k": (8, 4, 128, 4096, ()), "qwen16k": (4, 7, 128, 16384, (963, 963 * 2))}
ap = argparse.ArgumentParser()
ap.add_argument("--once", action="store_true")
ap.add_argument("--shape", nargs="*", default=list(SHAPES))
a = ap.parse_args()
dev = "cuda"
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("bf16_tflops", 1590.0) if (ROOT / "MEASURED_PEAKS.json").exists() else 1590.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    ts = []
    for i in range(iters + 3):
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return statistics.mean(ts), min(ts)


for name in a.shape:
    Hkv, G, D, S, wins = SHAPES[name]
    torch.manual_seed(0)
    # the [S, H, D] layout the qkv projection produces
    q = (torch.randn(1, S, Hkv * G, D, device=dev) * 1.3).to(torch.bfloat16).transpose(1, 2)
    k = torch.randn(1, S, Hkv, D, device=dev).to(torch.bfloat16).transpose(1, 2)
    v = torch.randn(1, S, Hkv, D, device=dev).to(torch.bfloat16).transpose(1, 2)
    fn = lambda: prefill_attention_stats(q, k, v, obs_len=16, windows=wins)  # noqa: E731
    fn(); torch.cuda.synchronize()
    if a.once:
        fn(); torch.cuda.synchronize()
        print(name, "launched")
        continue
    flops = 2.0 * S * S * D * Hkv * G
    ms, best = timed(fn)
    print(f"{name}: Hkv={Hkv} G={G} D={D} S={S} windows={wins}: {ms:.3f} ms (min {best:.3f}) -> "
          f"{flops / ms / 1e9:.1f} TFLOP/s algorithmic = {flops / ms / 1e9 / peak:.3f} of measured bf16 peak {peak:.0f}; "
          f"executed (3 GEMMs, full diagonal tiles) ~{1.5 * flops / ms / 1e9:.0f} TFLOP/s")
    sd = lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)  # noqa: E731
    sd(); ms2, _ = timed(sd)
    print(f"   torch F.scaled_dot_product_attention (y only, no probabilities): {ms2:.3f} ms")
    if S <= 8192:
        from oracle.prefill import blockwise_stats
        bw = lambda: blockwise_stats(q, k, v, obs_len=16, block=512, windows=wins)  # noqa: E731
        bw(); ms3, _ = timed(bw, iters=3)
        print(f"   torch restatement, block-wise (the previous prefill path): {ms3:.3f} ms")
