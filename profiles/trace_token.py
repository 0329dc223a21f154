#!/usr/bin/env python
"""Stage timeline of the whole-token kernel (cc_token_set_trace): for every stage of one decoded token and every
CTA, %globaltimer when the CTA entered the stage and when its input vector was staged.  Prints, per stage kind,
how long CTAs wait for their input (enter -> staged) and how long the stage's own work takes (staged -> next enter),
and the token's critical path.   python profiles/trace_token.py [--config 2] [--tokens 4]"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from cold_compress_b200 import generation_utils as gu  # noqa: E402
from cold_compress_b200.tokenizer import SyntheticTokenizer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=4)
ap.add_argument("--prompt", type=int, default=bench.PROMPT_LEN)
a = ap.parse_args()
dev = torch.device("cuda", 0)
model = bench.build_gpu_model(dev, 1)
gu.setup_caches(model, SyntheticTokenizer(), dev, bench.PROMPT_LEN + bench.MAX_NEW, bench.cache_kwargs())
torch.manual_seed(1234)
prompt = torch.randint(0, model.config.vocab_size, (a.prompt,), dtype=torch.int32, device=dev)
with torch.no_grad():
    tok, _ = gu.prefill(model, prompt.view(1, -1), torch.arange(a.prompt, device=dev))
    pos = torch.tensor([a.prompt], dtype=torch.int32, device=dev)
    tok = tok.view(1, 1).int()
    plan = model.token_plan()
    assert plan is not None, getattr(model, "token_kernel_reason", "")
    info = plan.info()
    print("plan:", info)
    n_ops, n_ctas = info["n_stages_program"], info["n_ctas"]
    trace = torch.zeros(n_ops, n_ctas, 12, dtype=torch.int64, device=dev)
    for i in range(a.tokens):
        t, _ = model.decode_token_fused(tok, pos)
        tok = t.view(1, 1).clone()
        pos += 1
    from cold_compress_b200 import _lib
    _lib.check(plan.lib.cc_token_set_trace(plan.handle, trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t, _ = model.decode_token_fused(tok, pos)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(plan.lib.cc_token_set_trace(plan.handle, None))
    print(f"token with tracing: {e0.elapsed_time(e1) * 1e3:.1f} us")
T = trace.cpu().numpy().astype(np.int64)
enter, staged = T[:, :, 0], T[:, :, 1]
t0 = enter[enter > 0].min()
kinds = []
L = len(model.layers)
for l in range(L):
    kinds += ["qkv", "attn", "wo", "gate", "w2"]
kinds += ["head"]
print(f"token span (first stage entered -> head staged, max over CTAs): {(staged[-1].max() - t0) / 1e3:.1f} us")
rows = {}
for oi, k in enumerate(kinds):
    act = enter[oi] > 0
    if k == "attn":
        act = staged[oi] > 0
    en, st = enter[oi][act], staged[oi][act]
    nxt = enter[oi + 1][act] if oi + 1 < len(kinds) else None
    wait = (st - en) / 1e3
    work = (nxt - st) / 1e3 if nxt is not None else None
    skew = (en.max() - en.min()) / 1e3
    rows.setdefault(k, []).append((np.median(wait), wait.max(), np.median(work) if work is not None else np.nan,
                                   work.max() if work is not None else np.nan, skew,
                                   (enter[oi + 1].max() - en.max()) / 1e3 if oi + 1 < len(kinds) else np.nan))
print(f"{'stage':>6} {'wait med':>9} {'wait max':>9} {'work med':>9} {'work max':>9} {'enter skew':>10} {'stage period':>12}   (us, mean over layers)")
for k, v in rows.items():
    m = np.nanmean(np.array(v), axis=0)
    print(f"{k:>6} {m[0]:9.2f} {m[1]:9.2f} {m[2]:9.2f} {m[3]:9.2f} {m[4]:10.2f} {m[5]:12.2f}")
# projection stages: when the warps of a CTA finish their rows, relative to "input staged"
print("\nwarp finish times after the input is staged (us; mean over layers of the per-stage median / max over CTAs and warps):")
for k in ("qkv", "wo", "gate", "w2"):
    med, mx, first = [], [], []
    for oi, kk in enumerate(kinds):
        if kk != k:
            continue
        d = (T[oi, :, 2:10] - staged[oi][:, None]) / 1e3
        d = np.where(T[oi, :, 2:10] > 0, d, np.nan)
        med.append(np.nanmedian(d)); mx.append(np.nanmax(d)); first.append(np.nanmin(d))
    print(f"{k:>6}: first warp {np.mean(first):6.2f}  median warp {np.mean(med):6.2f}  last warp {np.mean(mx):6.2f}")
# attention stage phases (active CTAs)
names = ["q/k/v arrived", "logits done", "X1 done", "exp + local sums", "X2 done", "probs", "P.V + CTA reduce", "state writes, X3 wait starts", "X3 done"]
print("\nattention stage phases, us after the stage was entered (mean over layers of the median / max over active CTAs):")
for si, nm in zip(range(1, 10), names):
    med, mx = [], []
    for oi, kk in enumerate(kinds):
        if kk != "attn":
            continue
        act = T[oi, :, 1] > 0
        d = (T[oi, act, si] - T[oi, act, 0]) / 1e3
        med.append(np.median(d)); mx.append(d.max())
    print(f"  {nm:>30}: {np.mean(med):6.2f}  {np.mean(mx):6.2f}")
# per-layer period
per_layer = [(enter[5 * (l + 1)].max() - enter[5 * l].max()) / 1e3 for l in range(L - 1)]
print(f"layer period: median {np.median(per_layer):.1f} us (roofline {436.2e6 / 6548.5e3:.1f} us of weights + K/V at the measured HBM peak)")
