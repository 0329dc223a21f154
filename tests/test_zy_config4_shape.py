"""BASELINE config 4's per-layer cache shape through the per-layer decode kernel (cc_decode_step), against the oracle.
(The file sorts late on purpose: the driver runs the suite with -x.)"""
import numpy as np
import pytest
import torch

from golden_utils import bits, check_y

pytestmark = pytest.mark.gpu


def test_hybrid_decode_at_the_baseline_config4_shape():
    """BASELINE config 4's per-layer shape: Qwen2-7B grouping (Hkv = 4, G = 7, D = 128), 17408 slots per head (17
    chunks per CTA through the 4-stage shared-memory ring), the 400-entry bf16 history ring, per-head FastGen
    strategies of bench.py's config 4 (window | window + heavy hitter | full).  Synthetic state: head 0 (window) at
    its budget -> evicts from the first step; head 1 (window + heavy hitter) at its budget -> evicts by ring score;
    head 2 (full) 16400 live rows -> appends; head 3 (window + heavy hitter) five tokens under budget -> appends, then
    evicts.  Eviction indices, per-head counts, positions and pooled probabilities bit-exact against the oracle."""
    from device_utils import make_device_cache, oracle_like
    Hkv, G, D, L, W = 4, 7, 128, 17408, 400
    strategies = [{"strategy": "window", "recent_window": 0.1176},
                  {"strategy": "window_heavy_hitter", "recent_window": 0.0588, "heavy_hitter_frac": 0.0588},
                  {"strategy": "full"}]
    cache = make_device_cache("hybrid", Hkv, D, L, hybrid_strategies=strategies)
    heads = [0, 1, 2, 1]
    cache.set_head_strategies(heads)
    g = int(cache.global_tokens)
    budget = [g + int(cache._hyb_rw[h]) + int(cache._hyb_hh[h]) for h in range(Hkv)]
    live = [budget[0], budget[1], 16400, budget[3] - 5]
    cur = 16400   # tokens seen so far
    gen = torch.Generator().manual_seed(417)
    dev = cache.k_cache.device
    for h in range(Hkv):
        n = live[h]
        cache.k_cache[0, h, :n] = (torch.randn(n, D, generator=gen) * 1.2).to(torch.bfloat16).to(dev)
        cache.v_cache[0, h, :n] = torch.randn(n, D, generator=gen).to(torch.bfloat16).to(dev)
        if n == cur:
            pos = torch.arange(cur, dtype=torch.int32)
        else:   # globals, a sorted random subset of the middle, the most recent tokens
            recent = n // 2
            mid = torch.sort(torch.randperm(cur - recent - g, generator=gen)[: n - recent - g]).values + g
            pos = torch.cat([torch.arange(g), mid, torch.arange(cur - recent, cur)]).int()
        cache.pos[0, h, :n] = pos.to(dev)
        cache.mask[0, h, 0, :n] = True
        cache.cache_cts[h] = n
        den = torch.randint(1, 3 * W, (n,), generator=gen, dtype=torch.int32)
        ring = (torch.rand(n, W, generator=gen) * (2.0 / n)).to(torch.bfloat16)
        ring[torch.arange(W).view(1, -1) >= den.view(-1, 1)] = 0   # a slot seen `den` times has that many ring entries
        cache.attn_history_num[0, h, :n] = ring.to(dev)
        cache.attn_history_denom[0, h, :n] = den.to(dev)
    cache.attn_counter.fill_(cur)
    cache.resync_history()
    cache._q_per_kv = G
    oc = oracle_like(cache, G)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    ids = torch.zeros(1, 1, dtype=torch.int32, device="cuda")
    y_means = []
    for t in range(12):
        q = (torch.randn(Hkv * G, D, generator=gen) * 1.5).to(torch.bfloat16)
        k = (torch.randn(Hkv, D, generator=gen) * 1.2).to(torch.bfloat16)
        v = torch.randn(Hkv, D, generator=gen).to(torch.bfloat16)
        ip.fill_(cur + t)
        y = cache.decode_step(q.view(1, Hkv * G, 1, D).cuda(), k.view(1, Hkv, 1, D).cuda(), v.view(1, Hkv, 1, D).cuda(),
                              ip, input_ids=ids)[0, :, 0].cpu()
        yo, idx_o, probs_o = oc.decode_step(q, k, v, cur + t, is_punc=False)
        assert torch.equal(cache._fill_idx.cpu(), idx_o), f"eviction index differs at step {t}: {cache._fill_idx.cpu()} vs {idx_o}"
        assert torch.equal(cache.cache_cts.cpu(), oc.cache_cts), f"per-head counts differ at step {t}"
        assert np.array_equal(bits(cache._out["probs"]), bits(probs_o)), f"pooled probabilities differ at step {t}"
        y_means.append(check_y(y, yo, G, (), f"step {t}"))
    assert np.mean(y_means) <= 1e-3
    assert torch.equal(cache.pos[0].cpu(), oc.pos) and torch.equal(cache.mask[0, :, 0].cpu(), oc.mask.view(Hkv, L))
    assert torch.equal(cache.attn_history_denom[0].cpu(), oc.attn_history_denom)
    assert np.array_equal(bits(cache.attn_history_num[0]), bits(oc.attn_history_num))
    assert cache.cache_cts.cpu().tolist() == [budget[0], budget[1], 16412, budget[3]]
    assert torch.equal(cache.attn_history_sum.cpu(), cache.attn_history_num.double().sum(-1).cpu())
