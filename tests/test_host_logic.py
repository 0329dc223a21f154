"""Host-side configuration logic against the reference's (CPU): cache-length normalisation, per-layer patterns
(tile / repeat / pyramid / funnel), per-layer recent windows, argument compatibility.  Expected values were produced
by the reference's own functions (generation_utils.py:229-321, cache.py:121-139); when the reference checkout is
present the functions are also compared live over a sweep."""
import argparse
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
PYRAMID_1024_8192_32 = [2036, 1916, 1852, 1796, 1732, 1668, 1604, 1548, 1484, 1420, 1364, 1300, 1236, 1172, 1116, 1052,
                        988, 932, 868, 804, 740, 684, 620, 556, 500, 436, 372, 308, 256, 256, 256, 256]


def test_cache_length_normalisation_and_patterns_recorded_values():
    from cold_compress_b200 import generation_utils as gu
    assert [gu.normalize_cache_length(*a) for a in ((64, 144), (0.5, 1000), (1.0, 17408), (1024, 8192), (0.1, 300))] == \
        [64, 504, 17408, 1024, 32]
    assert gu.apply_pattern([64, 128], 8, "tile", 8192) == [64, 64, 64, 64, 128, 128, 128, 128]
    assert gu.apply_pattern([64, 128], 8, "repeat", 8192) == [64, 128, 64, 128, 64, 128, 64, 128]
    assert gu.apply_pattern([1024], 8, "pyramid", 8192) == [2023, 1535, 1287, 1039, 791, 543, 295, 256]
    assert gu.apply_pattern([1024], 8, "funnel", 8192) == [256, 295, 543, 791, 1039, 1287, 1535, 2023]
    lens = gu.apply_pattern([1024], 32, "pyramid", 8192)   # BASELINE config 5
    assert lens == PYRAMID_1024_8192_32 and sum(lens) == 33128
    with pytest.raises(AssertionError):
        gu.apply_pattern([512, 1024], 32, "funnel", 8192)


@pytest.mark.skipif(not (REF / "generation_utils.py").exists(), reason="reference checkout not present")
def test_patterns_match_reference_over_a_sweep():
    from cold_compress_b200 import generation_utils as gu
    for m in ("generation_utils", "cache", "model", "prompt_compression", "tp", "tokenizer", "quantization_utils", "attention_utils"):
        sys.modules.pop(m, None)
    sys.path.insert(0, str(REF))
    try:
        import generation_utils as ref
        for seq in (144, 300, 1000, 4096, 8192, 17408):
            for v in (0.05, 0.1, 0.25, 0.5, 1.0, 16, 64, 250, 1024, 5000):
                assert gu.normalize_cache_length(v, seq) == ref.normalize_cache_length(v, seq), (v, seq)
        for n_layer in (8, 28, 32):
            for L in (256, 512, 1024, 2048):
                for pat in ("pyramid", "funnel"):
                    assert gu.apply_pattern([L], n_layer, pat, 8192) == ref.apply_pattern([L], n_layer, pat, 8192), (n_layer, L, pat)
            for pat in ("tile", "repeat"):
                if n_layer % 4 == 0:
                    assert gu.apply_pattern([1, 2, 3, 4], n_layer, pat, 8192) == ref.apply_pattern([1, 2, 3, 4], n_layer, pat, 8192)
    finally:
        sys.path.pop(0)
        for m in ("generation_utils", "cache", "model", "prompt_compression", "tp", "tokenizer", "quantization_utils", "attention_utils"):
            sys.modules.pop(m, None)


def test_setup_caches_expands_per_layer_values():
    """Per-layer lengths, strategies and integer recent windows as the reference derives them
    (generation_utils.py:324-388: recent_window <= 1 is a fraction of the layer's cache length)."""
    import torch
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.model import ModelArgs, Transformer
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    cfg = ModelArgs(block_size=2048, vocab_size=100, n_layer=4, n_head=4, dim=256, n_local_heads=2, intermediate_size=256,
                    rope_base=10000)
    model = Transformer(cfg).to(torch.bfloat16)
    ck = dict(max_cache_length=[0.25, 128], cache_bits=None, cache_length_pattern="repeat",
              cache_strategy=["recent_global", "heavy_hitter"], cache_strategy_pattern="repeat",
              prompt_compression_strategy=["recent_global", "heavy_hitter"], global_tokens=4, recent_window=0.25,
              history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)
    out = gu.setup_caches(model, SyntheticTokenizer(), torch.device("cpu"), 1000, ck)
    assert out["max_cache_length"] == [256, 128, 256, 128]
    assert out["cache_strategy"] == ["recent_global", "heavy_hitter"] * 2
    assert out["recent_window"] == [64, 32, 64, 32]
    kinds = [type(b.attention.kv_cache).__name__ for b in model.layers]
    assert kinds == ["KVCacheRecentGlobal", "KVCacheHeavyHitter"] * 2
    assert [b.attention.kv_cache.max_cache_length for b in model.layers] == [256, 128, 256, 128]
    assert model.layers[1].attention.kv_cache.recent_window == 32
    assert model.layers[1].attention.kv_cache.attn_history_num.dtype == torch.float64


def test_cache_argument_compatibility_checks():
    from cold_compress_b200.cache import add_cache_arguments, cache_compatibility
    p = argparse.ArgumentParser()
    add_cache_arguments(p)
    ok = p.parse_args(["--cache_strategy", "heavy_hitter", "--prompt_compression_strategy", "heavy_hitter",
                       "--max_cache_length", "0.5", "--cache_bits", "4", "--history_window_size", "8"])
    cache_compatibility(ok)
    assert ok.cache_bits == 4 and ok.history_window_size == 8
    for bad in (["--cache_strategy", "heavy_hitter", "--prompt_compression_strategy", "recent_global"],
                ["--cache_strategy", "full", "--max_cache_length", "0.5"],
                ["--cache_strategy", "hybrid", "--max_cache_length", "512", "--prompt_compression_strategy", "full"]):
        with pytest.raises(AssertionError):
            cache_compatibility(p.parse_args(bad))
