"""Hybrid (FastGen) head profiling at prefill, CPU: the S x S-free computation in
KVCacheHybrid.profile_and_update against a brute-force evaluation of the reference's definition
(masks per strategy over the full pooled attention map, cache.py:1066-1187), and -- when the reference
checkout is present -- against the reference's own profile_attn_heads."""
import math
import sys
from pathlib import Path

import pytest
import torch

HYBRID = [{"strategy": "window", "recent_window": 0.1},
          {"strategy": "window_heavy_hitter", "heavy_hitter_frac": 0.25, "recent_window": 0.1},
          {"strategy": "window_heavy_hitter", "heavy_hitter_frac": 0.5, "recent_window": 0.1},
          {"strategy": "full"}]
FASTGEN = [{"strategy": "special"}, {"strategy": "special_punc"},
           {"strategy": "special_punc_heavy_hitter", "heavy_hitter_frac": 0.3},
           {"strategy": "special_punc_heavy_hitter_window", "recent_window": 0.3, "heavy_hitter_frac": 0.3},
           {"strategy": "full"}]


def _setup(strategies, frac, S=96, L=160, H=4, G=2, D=64, seed=0, peaky=(0.3, 1.0, 3.0, 8.0)):
    from oracle.prefill import blockwise_stats as prefill_attention_stats
    from cold_compress_b200.cache import KVCacheHybrid
    torch.manual_seed(seed)
    q = torch.randn(1, H * G, S, D)
    for h, s in enumerate(peaky):
        q[:, h * G:(h + 1) * G] *= s
    q, k, v = q.to(torch.bfloat16), torch.randn(1, H, S, D).to(torch.bfloat16), torch.randn(1, H, S, D).to(torch.bfloat16)
    ids = torch.randint(0, 30, (1, S), dtype=torch.int32)
    cache = KVCacheHybrid(1, H, D, torch.bfloat16, max_cache_length=L, global_tokens=4, max_seq_length=L, cache_bits=None,
                          token_ids={"special": [[0], [1]], "punctuation": [5, 6, 7]}, min_recovery_frac=frac,
                          hybrid_strategies=strategies)
    return cache, q, k, v, ids


def _full_pooled_map(q, k):
    H, G = k.shape[1], q.shape[1] // k.shape[1]
    S, D = q.shape[2], q.shape[3]
    w = q @ k.repeat_interleave(G, dim=1).transpose(-2, -1) * (1 / math.sqrt(D))
    w = w + torch.zeros_like(w).masked_fill_(~torch.tril(torch.ones(S, S, dtype=torch.bool)), float("-inf"))
    return torch.softmax(w, dim=-1).view(1, H, G, S, S).mean(dim=2)


@pytest.mark.parametrize("strategies,frac", [(HYBRID, 0.4), (FASTGEN, 0.2), (FASTGEN, 0.8), (HYBRID, 0.95)])
def test_profile_matches_bruteforce_definition(strategies, frac):
    from oracle.prefill import blockwise_stats as prefill_attention_stats
    cache, q, k, v, ids = _setup(strategies, frac)
    S, H = q.shape[2], k.shape[1]
    pos = torch.arange(S)
    _, stats = prefill_attention_stats(q, k, v, windows=cache.profiling_windows(S), block=32)
    attn = _full_pooled_map(q, k)
    # the two reductions against the full map
    assert torch.allclose(stats.col_sum.float(), attn.sum(dim=2).float(), atol=3e-2, rtol=2e-2)
    cache.profile_and_update(pos, k, v, stats, input_ids=ids)
    # brute force: reference definition on the full map
    special = (ids[0] == 0) | (ids[0] == 1)
    punc = torch.isin(ids[0], torch.tensor([5, 6, 7], dtype=torch.int32))
    cum = (attn[0].sum(dim=1) / (S - pos)).float()
    idx = torch.arange(S)

    def masks(total_len):
        out = []
        for s in strategies:
            name = s["strategy"]
            m = torch.zeros(H, S, S, dtype=torch.bool)
            m[:, :, :4] = True
            if "special" in name:
                m |= special.view(1, 1, -1)
            if "punc" in name:
                m |= punc.view(1, 1, -1)
            if "window" in name:
                w = max(1, int(s["recent_window"] * total_len))
                m |= ((idx.view(-1, 1) - idx.view(1, -1)) < w) & (idx.view(1, -1) <= idx.view(-1, 1))
            if "heavy_hitter" in name:
                avail = (~m[0, -1]).nonzero().flatten()
                n = math.ceil(min(s["heavy_hitter_frac"] * total_len, avail.numel()))
                top = cum[:, avail].topk(n, dim=1).indices
                m.scatter_(2, avail[top].view(H, 1, n).expand(H, S, n), True)
            if name == "full":
                m.fill_(True)
            out.append(m)
        return torch.stack(out)

    score = attn.float().expand(len(strategies), -1, -1, -1).masked_fill(~masks(S), 0).sum(-1).mean(-1)
    want = (score >= frac).int().argmax(dim=0)
    borderline = ((score - frac).abs() < 2e-2).any(dim=0)  # bf16 row sums vs column/band sums
    got = cache.cache_strategies
    assert torch.equal(got[~borderline], want[~borderline]), (got, want, score)
    assert len(set(got.tolist())) > 1 or frac > 0.9, "test should exercise a mix of strategies"
    keep = masks(cache.max_cache_length)[got, torch.arange(H), -1]          # [H, S]
    for h in range(H):
        if borderline[h]:
            continue
        n = int(cache.cache_cts[h])
        assert n == int(keep[h].sum())
        assert cache.pos[0, h, :n].tolist() == keep[h].nonzero().flatten().tolist(), "kept tokens, ascending"
        assert torch.all(cache.pos[0, h, n:] == -1) and cache.mask[0, h, 0, :n].all() and not cache.mask[0, h, 0, n:].any()
        assert torch.equal(cache.k_cache[0, h, :n], k[0, h][keep[h]])
        if hasattr(cache, "punc_mask"):
            assert torch.equal(cache.punc_mask[0, h, :n], punc[keep[h]])
    if cache.requires_heavy_hitter:
        assert int(cache.attn_counter) == 1 and torch.all(cache.attn_history_denom == 1)
        assert torch.allclose(cache.attn_history_sum, cache.attn_history_num.double().sum(-1))


@pytest.mark.skipif(not Path("/root/reference/cache.py").exists(), reason="reference checkout not present")
def test_profile_matches_reference_choice():
    from oracle.prefill import blockwise_stats as prefill_attention_stats
    cache, q, k, v, ids = _setup(HYBRID, 0.6)
    S = q.shape[2]
    pos = torch.arange(S)
    _, stats = prefill_attention_stats(q, k, v, windows=cache.profiling_windows(S), block=48)
    cache.profile_and_update(pos, k, v, stats, input_ids=ids)
    for m in ("cache", "prompt_compression", "quantization_utils"):
        sys.modules.pop(m, None)
    sys.path.insert(0, "/root/reference")
    try:
        import cache as ref_cache
        ref = ref_cache.KVCacheHybrid(1, k.shape[1], k.shape[3], torch.bfloat16, max_cache_length=cache.max_cache_length,
                                      global_tokens=4, max_seq_length=cache.max_cache_length, cache_bits=None,
                                      token_ids={"special": [[0], [1]], "punctuation": [5, 6, 7]},
                                      min_recovery_frac=0.6, hybrid_strategies=HYBRID)
        ref.profile_and_update(pos, k, v, _full_pooled_map(q, k), input_ids=ids)
    finally:
        sys.path.pop(0)
        for m in ("cache", "prompt_compression", "quantization_utils"):
            sys.modules.pop(m, None)
    assert torch.equal(cache.cache_strategies, ref.cache_strategies.to(torch.int64))
    assert torch.equal(cache.cache_cts, ref.cache_cts.to(torch.int32))
    for h in range(k.shape[1]):  # same SET of kept tokens (the reference's order is an unstable argsort)
        n = int(cache.cache_cts[h])
        assert sorted(ref.pos[0, h, :n].tolist()) == cache.pos[0, h, :n].tolist()
