#!/usr/bin/env python
"""First-light diagnostics of the tcgen05 prefill kernel (csrc/prefill_attn.cu) against the torch restatement
(oracle/prefill.py, used here only as the checker).  Each case prints where the first disagreement is, so that
a descriptor / layout bug can be told apart from a numerics difference from one run:

    python profiles/debug_prefill.py [case ...]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cold_compress_b200.attention_utils import prefill_attention_stats  # noqa: E402
from oracle.prefill import blockwise_stats  # noqa: E402

dev = "cuda"


def report(name, got, want, tol):
    g, w = got.float().cpu(), want.float().cpu()
    d = (g - w).abs()
    bad = (d > tol * (1 + w.abs())).nonzero()
    print(f"  {name:10s} max|d|={float(d.max()):.4g}  mean|d|={float(d.mean()):.3g}  bad={bad.shape[0]}/{d.numel()}"
          + (f"  first bad idx={bad[0].tolist()} got={float(g[tuple(bad[0])]):.5g} want={float(w[tuple(bad[0])]):.5g}" if bad.shape[0] else ""))
    if bad.shape[0]:
        dims = [sorted(set(bad[:, i].tolist()))[:12] for i in range(bad.shape[1])]
        print(f"    bad index values per dim (first 12): {dims}")
    return bad.shape[0] == 0


def case(name, Hkv, G, D, S, vmode="rand", windows=(), qscale=1.0, seed=0):
    torch.manual_seed(seed)
    q = (torch.randn(1, Hkv * G, S, D) * qscale).to(torch.bfloat16)
    k = torch.randn(1, Hkv, S, D).to(torch.bfloat16)
    if vmode == "rand":
        v = torch.randn(1, Hkv, S, D).to(torch.bfloat16)
    elif vmode == "ones":
        v = torch.ones(1, Hkv, S, D).to(torch.bfloat16)
    elif vmode == "eye":  # v[j, d] = 1 if j % D == d: y[i, d] = sum of p over keys j = d (mod D)
        v = torch.zeros(1, Hkv, S, D)
        j = torch.arange(S)
        v[:, :, j, j % D] = 1.0
        v = v.to(torch.bfloat16)
    elif vmode == "keyidx":  # v[j, :] = j / 64: y = expected key index / 64
        v = (torch.arange(S).float() / 64).view(1, 1, S, 1).expand(1, Hkv, S, D).to(torch.bfloat16).contiguous()
    print(f"case {name}: Hkv={Hkv} G={G} D={D} S={S} v={vmode} windows={windows}")
    y_ref, st_ref = blockwise_stats(q, k, v, obs_len=16, block=64, windows=windows)
    y, st = prefill_attention_stats(q.to(dev), k.to(dev), v.to(dev), obs_len=16, windows=windows)
    torch.cuda.synchronize()
    ok = report("y", y[0], y_ref[0], 2e-2)
    ok &= report("col_sum", st.col_sum[0], st_ref.col_sum[0], 2e-2)
    ok &= report("obs_mean", st.obs_mean[0], st_ref.obs_mean[0], 2e-2)
    for w in windows:
        ok &= report(f"band{w}", st.band_sum[w], st_ref.band_sum[w], 2e-2)
    print("  ->", "OK" if ok else "MISMATCH")
    return ok


CASES = {
    "t1": lambda: case("one key tile", 1, 4, 128, 128),
    "t1ones": lambda: case("one key tile, V=1", 1, 4, 128, 128, vmode="ones"),
    "t1eye": lambda: case("one key tile, V=eye", 1, 4, 128, 128, vmode="eye"),
    "t1idx": lambda: case("one key tile, V=key index", 1, 4, 128, 128, vmode="keyidx"),
    "t3": lambda: case("three key tiles", 2, 4, 128, 320, windows=(5, 40)),
    "g7": lambda: case("G=7 ragged", 2, 7, 128, 203, windows=(21,)),
    "d64": lambda: case("D=64", 2, 2, 64, 150),
    "g1": lambda: case("G=1 (QB=128)", 2, 1, 128, 300),
    "short": lambda: case("S=9", 1, 4, 128, 9),
    "peaky": lambda: case("peaky logits", 2, 4, 128, 700, qscale=4.0, windows=(70,)),
    "long": lambda: case("S=2100", 4, 4, 128, 2100, windows=(210,)),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    res = {n: CASES[n]() for n in names}
    print({n: ("ok" if r else "FAIL") for n, r in res.items()})
    sys.exit(0 if all(res.values()) else 1)
