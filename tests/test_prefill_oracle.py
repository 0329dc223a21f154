"""CPU: oracle/prefill.py (restatement of the reference's prefill attention + pooled-map reductions) against
the vectors recorded from the unmodified reference (tests/golden/prefill_*.npz, make_golden_prefill.py)."""
import pytest
import torch

from golden_utils import check_y, y_error
from prefill_utils import PREFILL_FIXTURES, PrefillGolden, assert_close_bf16, keep_sets_equal_up_to_ties


@pytest.mark.parametrize("name", [n for n in PREFILL_FIXTURES if "math" not in n])
def test_full_map_restatement_matches_reference(name):
    from oracle.prefill import full_map, reductions
    g = PrefillGolden(name)
    q, k, v = g.qkv()
    y, pooled = full_map(q, k, v)
    check_y(y[0].reshape(-1, g.D), g.t["y"].reshape(-1, g.D), g.G * g.S, (), name)
    st = reductions(pooled, obs_len=16)
    assert_close_bf16(st.col_sum[0], g.t["col_sum"], f"{name} col_sum", min_equal=0.999)
    assert_close_bf16(st.obs_mean[0], g.t["obs_mean"], f"{name} obs_mean", min_equal=0.999)
    rows = g.t["rows_idx"].long()
    assert_close_bf16(pooled[0][:, rows, :], g.t["rows"], f"{name} pooled rows", min_equal=0.999)


@pytest.mark.parametrize("name", [n for n in PREFILL_FIXTURES if "math" not in n])
def test_blockwise_restatement_matches_reference(name):
    """The S x S-free formulation (what the CUDA kernel computes) against the reference's reductions."""
    from oracle.prefill import blockwise_stats
    g = PrefillGolden(name)
    q, k, v = g.qkv()
    y, st = blockwise_stats(q, k, v, obs_len=16, block=64, windows=(5, 40))
    check_y(y[0].reshape(-1, g.D), g.t["y"].reshape(-1, g.D), g.G * g.S, (), name)
    assert_close_bf16(st.col_sum[0], g.t["col_sum"], f"{name} col_sum", min_equal=0.99)
    assert_close_bf16(st.obs_mean[0], g.t["obs_mean"], f"{name} obs_mean", min_equal=0.99)
    # bands: a window wider than the prompt is the column sum, and bands are monotone in the window
    _, wide = blockwise_stats(q, k, v, block=64, windows=(g.S + 1,))
    assert torch.allclose(wide.band_sum[g.S + 1], st.col_sum[0].float(), rtol=1e-2, atol=1e-2)
    assert torch.all(st.band_sum[5] <= st.band_sum[40] + 1e-6)


@pytest.mark.parametrize("name", [n for n in PREFILL_FIXTURES if "math" not in n and "short" not in n])
def test_compressor_on_stats_matches_reference_keep_set(name):
    """Host logic: the SnapKV compressor fed with the two reductions picks the reference's tokens (up to ties
    at the cut-off) and seeds the same cumulative attention."""
    from cold_compress_b200.prompt_compression import PrefillAttnStats, get_prompt_compressor_constructor
    g = PrefillGolden(name)
    q, k, v = g.qkv()
    stats = PrefillAttnStats(col_sum=g.t["col_sum"].unsqueeze(0), obs_mean=g.t["obs_mean"].unsqueeze(0), obs_len=16)
    comp = get_prompt_compressor_constructor("heavy_hitter")(head_specific=True, max_cache_length=g.L, global_tokens=4,
                                                             recent_window=10)
    pos = torch.arange(g.S)
    keep, _, _, cum = comp(pos, k, v, attn=stats)
    assert torch.equal(comp._token_importances(pos, k, v, attn=stats)[0].float(), g.t["priority"].float())
    keep_sets_equal_up_to_ties(keep, g.t["keep_idxs"].long(), g.t["priority"], name)
    same = [h for h in range(g.Hkv) if torch.equal(keep[h], g.t["keep_idxs"][h].long())]
    assert same, "at least one head should have no tie at the cut-off in these fixtures"
    for h in same:
        assert torch.equal(cum[0, h].float(), g.t["cum_attn"][h].float())


def test_math_fixture_matches_sdpa_semantics():
    """l2 / recent_global prefill goes through F.scaled_dot_product_attention on bf16 tensors (model.py:402-411 with
    return_attn=False -> attention_utils.py:27-35).  The arithmetic is torch's fused kernel (un-vendored third party):
    the restatement is the same library call; against exact fp32 attention the REFERENCE's own output is ~2e-3
    normwise off (bf16 probabilities inside the fused kernel), which bounds what any bf16 P.V path can be held to."""
    g = PrefillGolden("prefill_l2_math")
    q, k, v = g.qkv()
    F = torch.nn.functional
    y = F.scaled_dot_product_attention(q, k.repeat_interleave(g.G, 1), v.repeat_interleave(g.G, 1), is_causal=True)
    assert_close_bf16(y[0], g.t["y"], "bf16 sdpa", max_ulp=2.0, min_equal=0.95)
    exact = F.scaled_dot_product_attention(q.float(), k.float().repeat_interleave(g.G, 1),
                                           v.float().repeat_interleave(g.G, 1), is_causal=True).to(torch.bfloat16)
    norm_rel, _ = y_error(g.t["y"].reshape(-1, g.D), exact[0].reshape(-1, g.D))
    assert 5e-4 < float(norm_rel.mean()) < 4e-3 and float(norm_rel.max()) < 2.0 ** -7
