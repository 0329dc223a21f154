import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path('/root/repo'); sys.path.insert(0, str(ROOT))
import bench
from cold_compress_b200 import generation_utils as gu, _lib
from cold_compress_b200.tokenizer import SyntheticTokenizer
dev = torch.device("cuda", 0)
model = bench.build_gpu_model(dev, 1)
gu.setup_caches(model, SyntheticTokenizer(), dev, bench.PROMPT_LEN + bench.MAX_NEW, bench.cache_kwargs())
torch.manual_seed(1234)
S = 2048
prompt = torch.randint(0, model.config.vocab_size, (S,), dtype=torch.int32, device=dev)
with torch.no_grad():
    tok, _ = gu.prefill(model, prompt.view(1, -1), torch.arange(S, device=dev))
    pos = torch.tensor([S], dtype=torch.int32, device=dev); tok = tok.view(1, 1).int()
    plan = model.token_plan(); info = plan.info(); n_ops, n_ctas = info["n_stages_program"], info["n_ctas"]
    main = n_ops * n_ctas * 12
    trace = torch.zeros(main + n_ops * 8 * 32, dtype=torch.int64, device=dev)
    for i in range(3):
        t, _ = model.decode_token_fused(tok, pos); tok = t.view(1, 1).clone(); pos += 1
    _lib.check(plan.lib.cc_token_set_trace(plan.handle, trace.data_ptr()))
    model.decode_token_fused(tok, pos); torch.cuda.synchronize()
T = trace.cpu().numpy()
dbg = T[main:].reshape(n_ops, 8, 32)
kinds = ["qkv", "attn", "wo", "gate", "w2"]
for oi in range(50, 55):
    k = kinds[oi % 5]
    if k == "attn": continue
    print(f"--- stage {oi} ({k}), CTA 0: per warp, cycles from loop start: tile i -> (wait done, consume done, release done)")
    for w in range(8):
        r = dbg[oi, w]
        if r[0] == 0: print(f" warp {w}: no tiles"); continue
        t0 = r[0]
        out = []
        for i in range(8):
            a, b, c = r[1 + 3 * i: 4 + 3 * i]
            if a == 0: break
            out.append(f"({a - t0},{b - t0},{c - t0})")
        print(f" warp {w}: " + " ".join(out))
