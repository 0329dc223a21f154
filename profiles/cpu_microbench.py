#!/usr/bin/env python
"""Why is the CPU baseline slow on a many-core host?  Times the two halves of one CPU decode
layer (torch bf16 GEMV, oracle cache step) at several thread counts."""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.oracle import OracleCache, lib  # noqa: E402

print("cpu_count", os.cpu_count(), "torch threads default", torch.get_num_threads(), torch.__config__.parallel_info().splitlines()[0:3])
w = (torch.randn(14336, 4096) * 0.02).to(torch.bfloat16)
x = torch.randn(1, 1, 4096).to(torch.bfloat16)
for nt in (1, 8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(nt)
    for _ in range(2):
        torch.nn.functional.linear(x, w)
    t0 = time.perf_counter()
    for _ in range(10):
        torch.nn.functional.linear(x, w)
    dt = (time.perf_counter() - t0) / 10
    print(f"torch bf16 GEMV 14336x4096 threads={nt}: {dt * 1e3:.2f} ms  ({w.numel() * 2 / dt / 1e9:.1f} GB/s)")
wf = w.float()
xf = x.float()
torch.set_num_threads(32)
t0 = time.perf_counter()
for _ in range(10):
    torch.nn.functional.linear(xf, wf)
print(f"torch fp32 GEMV threads=32: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")
oc = OracleCache("heavy_hitter", 8, 4, 128, 1024)
oc.k_cache.normal_(); oc.v_cache.normal_(); oc.mask.fill_(True); oc.pos.copy_(torch.arange(1024).expand(8, 1024)); oc.cache_cts.fill_(1024)
q = torch.randn(32, 128).to(torch.bfloat16); k = torch.randn(8, 128).to(torch.bfloat16); v = torch.randn(8, 128).to(torch.bfloat16)
for nt in (1, 8, 32, os.cpu_count()):
    os.environ["OMP_NUM_THREADS"] = str(nt)
    t0 = time.perf_counter()
    for i in range(20):
        oc.decode_step(q, k, v, 2000 + i)
    print(f"oracle decode step (8 heads, L=1024) OMP threads env={nt} (effective {lib().cc_oracle_threads()}): {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms")
