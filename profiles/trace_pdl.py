#!/usr/bin/env python
"""Do consecutive decode-step launches overlap under programmatic dependent launch?
Each of 6 back-to-back launches (6 different layers) gets its own %globaltimer trace buffer."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cold_compress_b200 import _lib  # noqa: E402
from device_utils import make_device_cache, synthetic_fill  # noqa: E402

Hkv, G, D, L = 8, 4, 128, 1024
caches = [make_device_cache("heavy_hitter", Hkv, D, L) for _ in range(6)]
for i, c in enumerate(caches):
    synthetic_fill(c, L, seed=i); c._q_per_kv = G
q = torch.randn(1, Hkv * G, 1, D, device="cuda").to(torch.bfloat16)
k = torch.randn(1, Hkv, 1, D, device="cuda").to(torch.bfloat16)
v = torch.randn(1, Hkv, 1, D, device="cuda").to(torch.bfloat16)
pos = torch.tensor([L], dtype=torch.int32, device="cuda")
for c in caches:
    c.decode_step(q, k, v, pos)
lib = _lib.load()
trace = torch.zeros(len(caches), 256, 16, dtype=torch.int64, device="cuda")
for overlap in (False, True):
    trace.zero_(); pos += 1
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(graph, stream=stream):
            for i, c in enumerate(caches):
                lib.cc_debug_set_trace(trace[i].data_ptr())
                c.decode_step(q, k, v, pos, overlap_predecessor=overlap)
        lib.cc_debug_set_trace(None)
        for _ in range(3):
            graph.replay()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().astype(np.int64)
    t0 = t[0, :128, 0].min()
    print(f"--- overlap_predecessor={overlap}")
    for i in range(len(caches)):
        s, w, e = t[i, :128, 0], t[i, :128, 2], t[i, :128, 13]
        print(f"  launch {i}: first CTA start {(s.min() - t0) / 1e3:7.2f} us, last CTA start {(s.max() - t0) / 1e3:7.2f}, "
              f"candidate done (median) {(np.median(w) - t0) / 1e3:7.2f}, end {(e.max() - t0) / 1e3:7.2f}")
