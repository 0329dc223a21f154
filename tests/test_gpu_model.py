"""GPU tests of the model-level decode path: fused one-token kernels vs tensor ops around the same
cache kernel, CUDA-graph replay vs eager, and generate() plumbing."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def small_model(strategy="heavy_hitter", L=64, n_layer=2, seed=0, device="cuda", bits=None):
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.model import ModelArgs, Transformer
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    cfg = ModelArgs(block_size=512, vocab_size=1000, n_layer=n_layer, n_head=8, dim=1024, n_local_heads=2,
                    intermediate_size=1024, rope_base=500000)
    torch.manual_seed(seed)
    model = Transformer(cfg)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "norm" not in name:
                p.normal_(0.0, 0.05)
    model = model.to(device=device, dtype=torch.bfloat16).eval()
    ck = dict(max_cache_length=[float(L)], cache_bits=bits, cache_length_pattern="tile", cache_strategy=[strategy],
              cache_strategy_pattern="tile", prompt_compression_strategy=[strategy.replace("debug_", "")],
              global_tokens=4, recent_window=10, history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)
    gu.setup_caches(model, SyntheticTokenizer(), torch.device(device), 400, ck)
    return model


@pytest.mark.parametrize("strategy", ["heavy_hitter", "l2", "recent_global"])
def test_fused_decode_matches_tensor_op_decode(strategy):
    from cold_compress_b200 import generation_utils as gu
    a = small_model(strategy)
    b = small_model(strategy)
    b.fused_decode = False
    torch.manual_seed(5)
    prompt = torch.randint(0, 1000, (150,), dtype=torch.int32, device="cuda")
    with torch.no_grad():
        ta, _ = gu.prefill(a, prompt.view(1, -1), torch.arange(150, device="cuda"))
        tb, _ = gu.prefill(b, prompt.view(1, -1), torch.arange(150, device="cuda"))
        assert torch.equal(ta, tb)
        for la, lb in zip(a.layers, b.layers):
            assert torch.equal(la.attention.kv_cache.pos, lb.attention.kv_cache.pos)
        pos = torch.tensor([150], dtype=torch.int32, device="cuda")
        tok = ta.view(1, 1)
        for t in range(12):
            la = a(tok, pos, is_prefill=False).float()
            lb = b(tok, pos, is_prefill=False).float()
            err = (la - lb).abs().max().item()
            scale = lb.abs().max().item()
            assert err <= 0.03 * scale + 0.03, f"step {t}: logits differ by {err} (scale {scale})"
            if t < 3:  # same eviction decisions while the two paths have not drifted apart
                for xa, xb in zip(a.layers, b.layers):
                    assert torch.equal(xa.attention.kv_cache._fill_idx, xb.attention.kv_cache._fill_idx)
            tok = torch.randint(0, 1000, (1, 1), dtype=torch.int32, device="cuda")
            pos += 1


def test_graph_replay_equals_eager_and_generate_runs():
    from cold_compress_b200 import generation_utils as gu
    a = small_model("heavy_hitter", seed=3)
    b = small_model("heavy_hitter", seed=3)
    torch.manual_seed(9)
    prompt = torch.randint(0, 1000, (120,), dtype=torch.int32, device="cuda")
    prefill, eager = gu.compile_funcs(False)
    _, graphed = gu.compile_funcs(True)
    seq_a, probs_a, stats = gu.generate(a, prompt, prefill, eager, max_new_tokens=24)
    seq_b, probs_b, _ = gu.generate(b, prompt, prefill, graphed, max_new_tokens=24)
    assert seq_a.numel() == 120 + 24
    assert torch.equal(seq_a, seq_b), "CUDA-graph decode and eager decode produced different tokens"
    for la, lb in zip(a.layers, b.layers):
        for n in ("pos", "k_cache", "attn_history_num", "attn_history_denom", "cache_cts"):
            assert torch.equal(getattr(la.attention.kv_cache, n), getattr(lb.attention.kv_cache, n)), n
    assert stats["decode_tokens"] == 24 and stats["decode_toks_per_sec"] > 0
    cs = a.get_cache_stats(120, 24)
    assert 0 < cs["compression_ratio_avg"] < 1


def test_ops_refuse_cpu_tensors():
    from cold_compress_b200 import _lib
    from device_utils import make_device_cache
    c = make_device_cache("heavy_hitter", 2, 128, 64, device="cpu")
    z = torch.zeros(1, 8, 1, 128, dtype=torch.bfloat16)
    with pytest.raises(_lib.CcError):
        c.decode_step(z, z[:, :2], z[:, :2], torch.zeros(1, dtype=torch.int32))


def test_hybrid_end_to_end_prefill_profile_and_graph_decode():
    """FastGen hybrid: prefill profiles heads (host tensor code), decode runs in the fused kernel with the
    device-side punctuation flag, eager and CUDA-graph decode agree token for token."""
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.model import ModelArgs, Transformer
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    strategies = [{"strategy": "special"}, {"strategy": "special_punc"},
                  {"strategy": "special_punc_heavy_hitter", "heavy_hitter_frac": 0.3},
                  {"strategy": "special_punc_heavy_hitter_window", "recent_window": 0.3, "heavy_hitter_frac": 0.3},
                  {"strategy": "full"}]

    def build():
        cfg = ModelArgs(block_size=512, vocab_size=64, n_layer=2, n_head=8, dim=1024, n_local_heads=2,
                        intermediate_size=1024, rope_base=500000)
        torch.manual_seed(4)
        m = Transformer(cfg)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if "norm" not in name:
                    p.normal_(0.0, 0.05)
        m = m.to(device="cuda", dtype=torch.bfloat16).eval()
        ck = dict(max_cache_length=[1.0], cache_bits=None, cache_length_pattern="tile", cache_strategy=["hybrid"],
                  cache_strategy_pattern="tile", prompt_compression_strategy=["full"], global_tokens=4,
                  recent_window=10, history_window_size=1, attn_thresholding=False, min_recovery_frac=0.5,
                  hybrid_strategies=strategies)
        gu.setup_caches(m, SyntheticTokenizer(vocab_size=64), torch.device("cuda"), 160, ck)
        return m

    a, b = build(), build()
    torch.manual_seed(2)
    prompt = torch.randint(0, 64, (100,), dtype=torch.int32, device="cuda")
    prefill, eager = gu.compile_funcs(False)
    _, graphed = gu.compile_funcs(True)
    seq_a, _, _ = gu.generate(a, prompt, prefill, eager, max_new_tokens=40)
    seq_b, _, _ = gu.generate(b, prompt, prefill, graphed, max_new_tokens=40)
    assert torch.equal(seq_a, seq_b)
    for la, lb in zip(a.layers, b.layers):
        ca, cb = la.attention.kv_cache, lb.attention.kv_cache
        assert torch.equal(ca.cache_cts, cb.cache_cts) and torch.equal(ca.pos, cb.pos)
        assert int(ca.cache_cts.max()) <= 100 + 40 and int(ca.cache_cts.min()) >= 4
    stats = a.get_cache_stats(100, 40)
    assert "compression_ratio_avg" in stats and "avg_strategy_idx_avg" in stats


def test_debug_cache_end_to_end_records_attention_loss():
    """debug_heavy_hitter through setup_caches / generate: full-cache attention, shadow heavy-hitter cache, attention
    loss per decode step; eager and CUDA-graph decode agree; a larger shadow cache loses less attention mass."""
    from cold_compress_b200 import generation_utils as gu
    torch.manual_seed(21)
    prompt = torch.randint(0, 1000, (150,), dtype=torch.int32, device="cuda")
    prefill, eager = gu.compile_funcs(False)
    losses = {}
    for L in (32, 96):
        _, graphed = gu.compile_funcs(True)   # a captured graph belongs to one model
        a = small_model("debug_heavy_hitter", L=L, seed=6)
        b = small_model("debug_heavy_hitter", L=L, seed=6)
        seq_a, _, _ = gu.generate(a, prompt, prefill, eager, max_new_tokens=30)
        seq_b, _, _ = gu.generate(b, prompt, prefill, graphed, max_new_tokens=30)
        assert torch.equal(seq_a, seq_b)
        ca, cb = a.layers[0].attention.kv_cache, b.layers[0].attention.kv_cache
        assert type(ca).__name__ == "KVCacheAnalysis" and ca.compressed.max_cache_length == L
        # generate() feeds the last sampled token back only as an output, so 29 decode steps touched the caches
        n = int(ca.attention_loss_ctr)
        assert n >= 29 and torch.equal(ca.attention_losses[:n], cb.attention_losses[:n])
        assert int(ca.compressed.cache_cts) == L and int((ca.pos >= 0).sum()) == 150 + n
        stats = a.get_cache_stats(150, 30)
        assert 0.0 <= stats["attention_loss_avg"] <= 1.0
        losses[L] = stats["attention_loss_avg"]
    assert losses[96] < losses[32], losses


@pytest.mark.parametrize("strategy,bits", [("heavy_hitter", 8), ("recent_global", 4)])
def test_quantised_cache_end_to_end(strategy, bits):
    """--cache_bits through setup_caches / generate: eager and CUDA-graph decode agree token for token, the cache the
    kernels read is exactly dequantize(codes), and the statistics report the reduced footprint."""
    from cold_compress_b200 import generation_utils as gu
    from oracle import quant
    torch.manual_seed(13)
    prompt = torch.randint(0, 1000, (120,), dtype=torch.int32, device="cuda")
    prefill, eager = gu.compile_funcs(False)
    _, graphed = gu.compile_funcs(True)
    a, b = small_model(strategy, seed=8, bits=bits), small_model(strategy, seed=8, bits=bits)
    seq_a, _, _ = gu.generate(a, prompt, prefill, eager, max_new_tokens=24)
    seq_b, _, _ = gu.generate(b, prompt, prefill, graphed, max_new_tokens=24)
    assert torch.equal(seq_a, seq_b)
    full = small_model(strategy, seed=8)
    seq_f, _, _ = gu.generate(full, prompt, prefill, eager, max_new_tokens=24)
    assert seq_f.numel() == seq_a.numel()
    for la, lb in zip(a.layers, b.layers):
        ca, cb = la.attention.kv_cache, lb.attention.kv_cache
        assert torch.equal(ca.k_cache, cb.k_cache) and torch.equal(ca.k_codes, cb.k_codes)
        H, L, D = ca.n_heads, ca.max_cache_length, ca.head_dim
        per = 8 // bits
        codes = ((ca.k_codes.cpu().to(torch.int32).view(-1, 1) >> torch.arange(0, 8, bits, dtype=torch.int32)) & (2 ** bits - 1))
        codes = codes.reshape(-1)[: H * L * D].view(H, L, D) if per > 1 else ca.k_codes.cpu().to(torch.int32).view(H, L, D)
        assert torch.equal(quant.dequantize(codes, ca.k_scales.cpu(), ca.k_zero_points.cpu(), bits), ca.k_cache[0].cpu())
    stats = a.get_cache_stats(120, 24)
    assert stats["compression_ratio_avg"] > full.get_cache_stats(120, 24)["compression_ratio_avg"]
