#!/usr/bin/env python
"""Drive ONLY the fused decode kernel (no model): N cache layers in a synthetic full state, S decode
steps each.  Used under ncu (profiles/README.md) and for quick timing:

    python profiles/run_k1.py [--strategy heavy_hitter] [--layers 32] [--steps 20] [--L 1024] [--time]
"""
import argparse
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from device_utils import make_device_cache, synthetic_fill  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--strategy", default="heavy_hitter")
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--L", type=int, default=1024)
ap.add_argument("--Hkv", type=int, default=8)
ap.add_argument("--G", type=int, default=4)
ap.add_argument("--D", type=int, default=128)
ap.add_argument("--time", action="store_true", help="CUDA-graph replay timing with an L2 flush between replays")
a = ap.parse_args()

dev = "cuda"
caches = [make_device_cache(a.strategy, a.Hkv, a.D, a.L) for _ in range(a.layers)]
for i, c in enumerate(caches):
    synthetic_fill(c, a.L, seed=i)
    c._q_per_kv = a.G
gen = torch.Generator(device=dev).manual_seed(0)
q = (torch.randn(1, a.Hkv * a.G, 1, a.D, device=dev, generator=gen) * 1.5).to(torch.bfloat16)
k = torch.randn(1, a.Hkv, 1, a.D, device=dev, generator=gen).to(torch.bfloat16)
v = torch.randn(1, a.Hkv, 1, a.D, device=dev, generator=gen).to(torch.bfloat16)
pos = torch.tensor([a.L], dtype=torch.int32, device=dev)
if not a.time:
    for s in range(a.steps):
        for c in caches:
            c.decode_step(q, k, v, pos)
        pos += 1
    torch.cuda.synchronize()
    print("done", a.steps * a.layers, "launches")
else:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for c in caches:
            c.decode_step(q, k, v, pos)
        stream.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            for c in caches:
                c.decode_step(q, k, v, pos, overlap_predecessor=len(caches) > 1)
        ts = []
        for i in range(a.steps + 3):
            flush.fill_(i & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            graph.replay()
            e1.record(stream)
            e1.synchronize()
            pos += 1
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3 / len(caches))
    byts = 2 * a.Hkv * a.L * a.D * 2
    us = statistics.mean(ts)
    print(f"{a.strategy} L={a.L} Hkv={a.Hkv}: {us:.2f} us/launch (min {min(ts):.2f}), {byts / us / 1e3:.1f} GB/s algorithmic")
