#!/usr/bin/env python
"""Phase timeline of the fused decode kernel from in-kernel %globaltimer stamps (cc_debug_set_trace).
Prints, for the last of several back-to-back launches, the per-phase time (median / max over CTAs)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cold_compress_b200 import _lib  # noqa: E402
from device_utils import make_device_cache, synthetic_fill  # noqa: E402

strategy = sys.argv[1] if len(sys.argv) > 1 else "heavy_hitter"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
Hkv, G, D = 8, 4, 128
caches = [make_device_cache(strategy, Hkv, D, L) for _ in range(8)]
for i, c in enumerate(caches):
    synthetic_fill(c, L, seed=i); c._q_per_kv = G
q = torch.randn(1, Hkv * G, 1, D, device="cuda").to(torch.bfloat16)
k = torch.randn(1, Hkv, 1, D, device="cuda").to(torch.bfloat16)
v = torch.randn(1, Hkv, 1, D, device="cuda").to(torch.bfloat16)
pos = torch.tensor([L], dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for c in caches:
    c.decode_step(q, k, v, pos)
trace = torch.zeros(512, 16, dtype=torch.int64, device="cuda")
names = ["start", "tma issued", "candidate", "K chunk0 landed", "logits+lmax", "X1 done", "winner+gmax",
         "exp+local sums", "X2 done", "probs", "PV loop", "reduce+push y", "state writes+X3 done", "end"]
lib = _lib.load()
for rep in range(3):
    flush.fill_(rep)
    lib.cc_debug_set_trace(trace.data_ptr())
    pos += 1
    torch.cuda.synchronize()
    for c in caches:   # 8 back-to-back launches; the trace keeps the last one
        c.decode_step(q, k, v, pos)
    torch.cuda.synchronize()
    lib.cc_debug_set_trace(None)
    t = trace.cpu().numpy().astype(np.int64)
    n = int((t[:, 0] > 0).sum())
    t = t[:n]
    t0 = t[:, 0].min()
    print(f"--- {strategy} L={L} rep {rep}: {n} CTAs, kernel span {(t[:, 13].max() - t0) / 1e3:.2f} us, "
          f"start skew {(t[:, 0].max() - t0) / 1e3:.2f} us")
    for i in range(1, 14):
        if (t[:, i] == 0).all():
            continue
        d = (t[:, i] - t0) / 1e3
        print(f"  {names[i]:>20}: reached at median {np.median(d):6.2f} us  max {d.max():6.2f} us")
