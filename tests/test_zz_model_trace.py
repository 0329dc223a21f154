"""Whole-model parity against the unmodified reference (BASELINE config 1): Llama-3-8B shape,
32 layers, recent_global with 64 slots per layer, 128-token prompt, 16 generated tokens.  tests/golden/model_cfg1.npz
was recorded from the reference's own generate() on CPU by tests/golden/make_golden_model.py; both sides build the
same weights from tests/model_weights.py.  Checked on the GPU, through this package's prefill and the whole-token
decode kernel, with the reference's tokens fed back (so every step is compared, not only the steps before the first
legitimate flip):
  * the slot every layer overwrites at every decode step, and every layer's final `pos`: identical;
  * the token chosen at every step: a random-init model over a 128k vocabulary is almost flat at the top (the
    reference's own top-2 bf16 probabilities tie in several steps) and 32 layers of bf16 roundings in a different
    summation order move the top probabilities by up to ~20 % (measured on the B200 for the prefill step: 0.21), so
    our arg-max must be one of the reference's top-64 tokens with a reference probability of at least 0.6 of the
    reference's maximum, and our probabilities of the reference's top-64 tokens must agree with the reference's to
    35 % (a dropped row, a wrong rotation or a wrong residual scrambles the top-64 set altogether).
(The file sorts last on purpose: the driver runs the suite with -x.)"""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "model_cfg1.npz"


@pytest.mark.parametrize("token_kernel", [True, False])
def test_config1_whole_model_trace_matches_reference(token_kernel):
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.model import ModelArgs, Transformer
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    from model_weights import seeded_weights
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs a 16 GB model on the device")
    z = np.load(GOLD)
    prompt = torch.from_numpy(z["prompt"]).cuda()
    seq, fill_ref, pos_ref, top_p, top_i = z["seq"], z["fill_idx"], z["final_pos"], z["top_p"], z["top_i"]
    P, T = prompt.numel(), seq.size - prompt.numel()
    assert (seq[:P] == z["prompt"]).all() and T == 16 and fill_ref.shape == (T - 1, 32)
    cfg = ModelArgs.from_name("Meta-Llama-3-8B-Instruct")
    with torch.device("cuda"):
        model = Transformer(cfg).to(torch.bfloat16)
    seeded_weights(model).eval()
    model.token_kernel = token_kernel
    ck = dict(max_cache_length=[64.0], cache_bits=None, cache_length_pattern="tile", cache_strategy=["recent_global"],
              cache_strategy_pattern="tile", feed_long_prompts=False, prompt_compression_strategy=["recent_global"],
              global_tokens=4, recent_window=10, history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)
    gu.setup_caches(model, SyntheticTokenizer(vocab_size=cfg.vocab_size), torch.device("cuda"), P + T, ck)

    def check_token(t, tok, probs):
        ids, ref = top_i[t], top_p[t]
        tok = int(tok)
        assert tok in ids.tolist(), f"step {t}: token {tok} is not among the reference's 64 most likely tokens"
        assert ref[ids.tolist().index(tok)] >= 0.6 * ref[0], f"step {t}: token {tok} has reference probability {ref[ids.tolist().index(tok)]} < max {ref[0]}"
        ours = probs.float().view(-1)[torch.from_numpy(ids.astype(np.int64)).cuda()].cpu().numpy()
        rel = np.abs(ours - ref) / ref
        assert rel.max() <= 0.35, f"step {t}: probabilities of the reference's top tokens differ by {rel.max():.3f}"

    with torch.no_grad():
        tok, probs = gu.prefill(model, prompt.view(1, -1), torch.arange(P, device="cuda"))
        check_token(0, tok.view(-1)[0], probs)
        pos = torch.tensor([P], dtype=torch.int32, device="cuda")
        for t in range(1, T):
            x = torch.tensor([[int(seq[P + t - 1])]], dtype=torch.int32, device="cuda")   # the reference's token
            tok, probs = gu.decode_one_token(model, x, pos)
            assert (model._tplan is not None) == token_kernel
            got = [int(b.attention.kv_cache._fill_idx.view(-1)[0]) for b in model.layers]
            assert got == fill_ref[t - 1].tolist(), f"decode step {t}: layers overwrote {got}, the reference {fill_ref[t - 1].tolist()}"
            check_token(t, tok.view(-1)[0], probs)
            pos += 1
    for i, b in enumerate(model.layers):
        assert torch.equal(b.attention.kv_cache.pos.view(-1).cpu(), torch.from_numpy(pos_ref[i])), f"layer {i}: final pos differs"
