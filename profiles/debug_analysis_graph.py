#!/usr/bin/env python
"""Where do eager and CUDA-graph decode of a debug_heavy_hitter model first diverge?  (diagnostic)"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_model import small_model  # noqa: E402
from cold_compress_b200 import generation_utils as gu  # noqa: E402

torch.manual_seed(21)
prompt = torch.randint(0, 1000, (150,), dtype=torch.int32, device="cuda")
a = small_model("debug_heavy_hitter", L=32, seed=6)
b = small_model("debug_heavy_hitter", L=32, seed=6)
with torch.no_grad():
    ta, _ = gu.prefill(a, prompt.view(1, -1), torch.arange(150, device="cuda"))
    tb, _ = gu.prefill(b, prompt.view(1, -1), torch.arange(150, device="cuda"))
    print("prefill tokens equal:", torch.equal(ta, tb))
    dec = gu.GraphedDecoder()
    pa = torch.tensor([150], dtype=torch.int32, device="cuda"); pb = pa.clone()
    tok_a, tok_b = ta.view(1, 1).clone(), tb.view(1, 1).clone()
    for t in range(30):
        na, _ = gu.decode_one_token(a, tok_a, pa)
        nb, _ = dec(b, tok_b, pb)
        torch.cuda.synchronize()
        bad = []
        for li, (la, lb) in enumerate(zip(a.layers, b.layers)):
            ca, cb = la.attention.kv_cache, lb.attention.kv_cache
            for n in ("pos", "k_cache", "attention_losses", "attention_loss_ctr"):
                if not torch.equal(getattr(ca, n), getattr(cb, n)): bad.append((li, n))
            for n in ("pos", "k_cache", "attn_history_num", "attn_history_denom", "cache_cts"):
                if not torch.equal(getattr(ca.compressed, n), getattr(cb.compressed, n)): bad.append((li, "compressed." + n))
            if not torch.equal(ca._out["probs"], cb._out["probs"]): bad.append((li, "probs"))
            if not torch.equal(ca._out["y"], cb._out["y"]): bad.append((li, "y"))
        print(f"step {t}: tok eager {int(na)} graph {int(nb)} diverged: {bad[:8]}")
        if bad:
            li, n = bad[0]
            ca, cb = a.layers[li].attention.kv_cache, b.layers[li].attention.kv_cache
            x, y = (getattr(ca, n) if "." not in n and n not in ("probs", "y") else
                    (ca._out[n] if n in ("probs", "y") else getattr(ca.compressed, n.split(".")[1]))), None
            break
        tok_a, tok_b = na.view(1, 1).clone(), nb.view(1, 1).clone()
        pa += 1; pb += 1
