"""GPU: the tcgen05 prefill attention kernel (cc_prefill_attention through the product API
cold_compress_b200.attention_utils.prefill_attention_stats) against the vectors recorded from the unmodified
reference (tests/golden/prefill_*.npz), against the restatement in oracle/prefill.py, and -- at BASELINE sizes --
through size-independent properties."""
import pytest
import torch

from prefill_utils import PREFILL_FIXTURES, PrefillGolden, assert_close_bf16, check_prefill_y, keep_sets_equal_up_to_ties

pytestmark = pytest.mark.gpu
NAIVE_FIXTURES = [n for n in PREFILL_FIXTURES if "math" not in n]


def _run(q, k, v, **kw):
    from cold_compress_b200.attention_utils import prefill_attention_stats
    y, st = prefill_attention_stats(q.cuda(), k.cuda(), v.cuda(), **kw)
    torch.cuda.synchronize()
    return y, st


@pytest.mark.parametrize("name", NAIVE_FIXTURES)
def test_prefill_kernel_matches_reference_vectors(name):
    g = PrefillGolden(name)
    q, k, v = g.qkv()
    y, st = _run(q, k, v, obs_len=16)
    check_prefill_y(y[0].reshape(-1, g.D), g.t["y"].reshape(-1, g.D), name)
    # reductions of the pooled map: the same fp32 sums in another order, over probabilities that may differ by an
    # ulp where a logit sits on a bf16 rounding boundary -> equal or adjacent bf16 values
    assert_close_bf16(st.col_sum[0], g.t["col_sum"], f"{name} col_sum", max_ulp=1.0, min_equal=0.90, floor=2.0 ** -14)
    assert_close_bf16(st.obs_mean[0], g.t["obs_mean"], f"{name} obs_mean", max_ulp=1.0, min_equal=0.90, floor=2.0 ** -14)
    # one full row of pooled probabilities: an observation window of ONE row is the last row of the map
    _, st1 = _run(q, k, v, obs_len=1)
    last = g.t["rows"][:, -1, :]
    assert int(g.t["rows_idx"][-1]) == g.S - 1
    assert_close_bf16(st1.obs_mean[0], last, f"{name} pooled last row", max_ulp=1.0, min_equal=0.97, floor=2.0 ** -20)


@pytest.mark.parametrize("name", [n for n in NAIVE_FIXTURES if "short" not in n])
def test_prompt_compression_on_kernel_stats(name):
    """SnapKV compressor on the kernel's reductions: the reference's kept tokens (up to ties at the cut-off)."""
    from cold_compress_b200.prompt_compression import get_prompt_compressor_constructor
    g = PrefillGolden(name)
    q, k, v = g.qkv()
    _, st = _run(q, k, v, obs_len=16)
    comp = get_prompt_compressor_constructor("heavy_hitter")(head_specific=True, max_cache_length=g.L, global_tokens=4,
                                                             recent_window=10)
    pos = torch.arange(g.S, device="cuda")
    keep, _, _, cum = comp(pos, k.cuda(), v.cuda(), attn=st)
    keep_sets_equal_up_to_ties(keep.cpu(), g.t["keep_idxs"].long(), g.t["priority"], name)
    for h in range(g.Hkv):
        if torch.equal(keep[h].cpu(), g.t["keep_idxs"][h].long()):
            assert_close_bf16(cum[0, h], g.t["cum_attn"][h], f"{name} cum_attn head {h}", max_ulp=1.0, min_equal=0.85,
                              floor=2.0 ** -14)


@pytest.mark.parametrize("name", NAIVE_FIXTURES)
def test_prefill_kernel_matches_oracle_bands_and_layouts(name):
    from oracle.prefill import blockwise_stats
    g = PrefillGolden(name)
    q, k, v = g.qkv()
    wins = (3, max(4, g.S // 5), g.S + 7)
    y, st = _run(q, k, v, obs_len=16, windows=wins)
    _, ref = blockwise_stats(q, k, v, obs_len=16, block=64, windows=wins)
    for w in wins:
        assert torch.allclose(st.band_sum[w].cpu(), ref.band_sum[w], rtol=1e-2, atol=2e-3), f"{name} band {w}"
    # a window wider than the prompt IS the column sum (same additions in the same order)
    _, st0 = _run(q, k, v, obs_len=16)
    assert torch.equal(st.band_sum[g.S + 7].to(torch.bfloat16), st0.col_sum[0])
    # the [S, H, D] layout the projections produce (strided heads) gives the same bits as contiguous heads
    qs = q.cuda().transpose(1, 2).contiguous().transpose(1, 2)
    ks = k.cuda().transpose(1, 2).contiguous().transpose(1, 2)
    vs = v.cuda().transpose(1, 2).contiguous().transpose(1, 2)
    y2, st2 = _run(qs, ks, vs, obs_len=16)
    assert torch.equal(y2, y) and torch.equal(st2.col_sum, st0.col_sum) and torch.equal(st2.obs_mean, st0.obs_mean)
    # bit-reproducible run to run (fixed-order reduction, no floating-point atomics)
    y3, st3 = _run(q, k, v, obs_len=16, windows=wins)
    assert torch.equal(y3, y) and torch.equal(st3.col_sum, st.col_sum)
    assert all(torch.equal(st3.band_sum[w], st.band_sum[w]) for w in wins)


@pytest.mark.parametrize("Hkv,G,D,S,W", [(8, 4, 128, 8192, 819), (4, 7, 128, 16384, 1638), (2, 2, 64, 4096, 0)])
def test_prefill_full_size_properties(Hkv, G, D, S, W):
    """BASELINE-size prompts (configs 2-4 head shapes): rows of the pooled map sum to 1, so the column sums add up
    to S and the observation-window sums to obs_len; y is a convex combination of the values; y and the
    reductions agree with the torch restatement evaluated on the same device (checker only)."""
    from oracle.prefill import blockwise_stats
    torch.manual_seed(S + G)
    q = (torch.randn(1, Hkv * G, S, D, device="cuda") * 1.3).to(torch.bfloat16)
    k = torch.randn(1, Hkv, S, D, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, Hkv, S, D, device="cuda").to(torch.bfloat16)
    wins = (W,) if W else ()
    y, st = _run(q, k, v, obs_len=16, windows=wins)
    col = st.col_sum[0].float()
    assert torch.all(torch.isfinite(y.float())) and torch.all(col >= 0)
    assert torch.allclose(col.sum(dim=1), torch.full((Hkv,), float(S), device="cuda"), rtol=4e-3)
    assert torch.allclose(st.obs_mean[0].float().sum(dim=1), torch.ones(Hkv, device="cuda"), rtol=1e-2)
    vmax = v.float().abs().amax(dim=2, keepdim=True).repeat_interleave(G, dim=1)
    assert torch.all(y.float().abs() <= vmax * 1.01 + 1e-3)
    if W:
        assert torch.all(st.band_sum[W] <= col * 1.01 + 1e-2)
    y_ref, ref = blockwise_stats(q, k, v, obs_len=16, block=512, windows=wins)
    check_prefill_y(y[0].reshape(-1, D), y_ref[0].reshape(-1, D), f"S={S}")
    assert_close_bf16(st.col_sum[0], ref.col_sum[0], "col_sum", max_ulp=1.0, min_equal=0.85, floor=2.0 ** -14)
    assert_close_bf16(st.obs_mean[0], ref.obs_mean[0], "obs_mean", max_ulp=1.0, min_equal=0.85, floor=2.0 ** -14)
    if W:
        assert torch.allclose(st.band_sum[W], ref.band_sum[W], rtol=1e-2, atol=2e-3)


def test_prefill_argument_errors():
    from cold_compress_b200 import _lib, ops
    q = torch.zeros(8, 64, 128, dtype=torch.bfloat16, device="cuda")
    kv = torch.zeros(2, 64, 128, dtype=torch.bfloat16, device="cuda")
    y = torch.empty_like(q)
    with pytest.raises(_lib.CcError):   # statistics without a workspace
        ops.prefill_attention(q, kv, kv, y, _lib.ATTN_NAIVE_BF16, 16, [], torch.empty(2, 64, device="cuda"), None, None, None)
    with pytest.raises(_lib.CcError):   # head_dim 96 is not built
        q96 = torch.zeros(8, 64, 96, dtype=torch.bfloat16, device="cuda")
        ops.prefill_attention(q96, kv[..., :96].contiguous(), kv[..., :96].contiguous(), torch.empty_like(q96),
                              _lib.ATTN_NAIVE_BF16, 0, [], None, None, None, None)
    with pytest.raises(_lib.CcError):   # CPU tensors: no fallback
        ops.prefill_attention(q.cpu(), kv.cpu(), kv.cpu(), y.cpu(), _lib.ATTN_NAIVE_BF16, 0, [], None, None, None, None)
