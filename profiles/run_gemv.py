#!/usr/bin/env python
"""Achieved HBM bandwidth of the one-token GEMV kernels on the Llama-3-8B layer shapes.
32 distinct weight sets are cycled (1.7-7.5 GB per shape >> L2) inside one CUDA graph per shape."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cold_compress_b200 import ops  # noqa: E402

dev = "cuda"
dim, inter, Hq, Hkv, D, NL = 4096, 14336, 32, 8, 128, 32
bf = torch.bfloat16
mk = lambda *s: (torch.randn(*s, device=dev, dtype=torch.float32) * 0.02).to(bf)  # noqa: E731
x = mk(dim); nw = torch.ones(dim, device=dev, dtype=bf); res = mk(dim)
gate_in = mk(inter)
freqs = torch.randn(16384, D // 2, 2, device=dev).to(bf)
pos = torch.tensor([8000], dtype=torch.int32, device=dev)
q, k, v = mk(Hq * D), mk(Hkv * D), mk(Hkv * D)
out_d, out_i = mk(dim), mk(inter)
shapes = {
    "rmsnorm_qkv_rope 6144x4096": ([mk(dim + 2 * Hkv * D, dim) for _ in range(NL)],
                                   lambda w: ops.rmsnorm_qkv_rope(x, nw, 1e-5, w, None, freqs, pos, q, k, v, Hq, Hkv, D, [], [], None)),
    "gemv_residual(wo) 4096x4096": ([mk(dim, dim) for _ in range(NL)],
                                    lambda w: ops.gemv_residual(x, None, 0.0, w, res, out_d, [], [], None)),
    "rmsnorm_ffn_gate 2x14336x4096": ([(mk(inter, dim), mk(inter, dim)) for _ in range(NL)],
                                      lambda w: ops.rmsnorm_ffn_gate(x, nw, 1e-5, w[0], w[1], out_i, [], [], None)),
    "gemv_residual(w2) 4096x14336": ([mk(dim, inter) for _ in range(NL)],
                                     lambda w: ops.gemv_residual(gate_in, None, 0.0, w, res, out_d, [], [], None)),
}
tot_us = 0.0
for name, (ws, fn) in shapes.items():
    nbytes = sum(t.numel() * 2 for t in (ws[0] if isinstance(ws[0], tuple) else (ws[0],)))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for w in ws[:2]:
            fn(w)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for w in ws:
                fn(w)
        ts = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); g.replay(); e1.record(stream); e1.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3 / NL)
    us = statistics.mean(ts)
    tot_us += us
    print(f"{name:34s} {us:8.2f} us/launch  {nbytes / us / 1e3:8.1f} GB/s")
    del ws
    torch.cuda.empty_cache()
print(f"sum per layer {tot_us:.1f} us -> {32 * tot_us / 1e3:.2f} ms per token for the 4 layer GEMVs")
# reference point: cuBLAS via torch for the same shapes
for name, (n, kk) in {"torch F.linear 14336x4096": (inter, dim), "torch F.linear 4096x14336": (dim, inter),
                      "torch F.linear 6144x4096": (6144, dim)}.items():
    ws = [mk(n, kk) for _ in range(NL)]
    xin = mk(1, kk)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        torch.nn.functional.linear(xin, ws[0]); stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for w in ws:
                torch.nn.functional.linear(xin, w)
        ts = []
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); g.replay(); e1.record(stream); e1.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3 / NL)
    us = statistics.mean(ts)
    print(f"{name:34s} {us:8.2f} us/launch  {n * kk * 2 / us / 1e3:8.1f} GB/s  (cuBLAS reference point)")
    del ws
