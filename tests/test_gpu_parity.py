"""GPU parity tests (run on the B200 with `-m gpu`): the CUDA decode step, called through the
product API (cache classes -> torch custom op -> C ABI), against

  * the committed golden traces recorded from the unmodified reference, and
  * the CPU oracle run in lockstep on the same inputs.

Bars: eviction indices / pos / mask / counters / denominators bit-exact; pooled probabilities
bit-exact against the oracle (same arithmetic definition) and within the golden bar of
test_oracle_golden.py against the reference; y per golden_utils.check_y.
"""
import numpy as np
import pytest
import torch

from golden_utils import GPU_FIXTURES as ALL_FIXTURES, Golden, bits, check_y

pytestmark = pytest.mark.gpu

NON_HYBRID = [n for n in ALL_FIXTURES if not n.startswith("hybrid")]
HYBRID = [n for n in ALL_FIXTURES if n.startswith("hybrid")]


def _run_trace(name, **step_kwargs):
    from device_utils import device_cache_from_golden, oracle_from_golden
    g = Golden(name)
    cache = device_cache_from_golden(g)
    oc = oracle_from_golden(g)
    dev = "cuda"
    ip = torch.zeros(1, dtype=torch.int32, device=dev)
    prob_mismatch_ref = prob_total = 0
    y_means, y_means_ref = [], []
    for t in range(g.T):
        s = g.step(t)
        q = s["q"].view(1, g.Hkv * g.G, 1, g.D).to(dev)
        k = s["k_new"].view(1, g.Hkv, 1, g.D).to(dev)
        v = s["v_new"].view(1, g.Hkv, 1, g.D).to(dev)
        ip.fill_(s["input_pos"])
        y = cache.decode_step(q, k, v, ip, is_punc=s["is_punc"], **step_kwargs)
        y = y[0, :, 0].cpu()
        yo, idx_o, probs_o = oc.decode_step(s["q"], s["k_new"], s["v_new"], s["input_pos"],
                                            fill_idx=s["fill_idx"] if g.policy == "external" else None,
                                            is_punc=s["is_punc"])
        ref_idx = s["fill_idx"].numpy()
        got_idx = cache._fill_idx.cpu().numpy()
        want = np.broadcast_to(ref_idx, (g.Hkv,)) if ref_idx.size == 1 else ref_idx
        assert np.array_equal(got_idx, want), f"{name}: eviction index differs from the reference at step {t}"
        if g.policy != "external":
            assert np.array_equal(got_idx, idx_o.numpy()), f"{name}: eviction index differs from the oracle at step {t}"
        pos = cache.pos[0].cpu().numpy()
        assert np.array_equal(pos.astype(np.int16), s["pos_after"]), f"{name}: pos differs at step {t}"
        flipped = set()
        if probs_o is not None:
            got = bits(cache._out["probs"])
            assert np.array_equal(got, bits(probs_o)), f"{name}: pooled probabilities differ from the oracle at step {t}"
            diff = np.abs(got.astype(np.int32) - bits(s["probs"]).astype(np.int32))
            assert diff.max() <= 2
            prob_mismatch_ref += int((diff != 0).sum())
            prob_total += diff.size
            flipped = set(np.nonzero(diff.reshape(g.Hkv, -1).any(axis=1))[0].tolist())
        y_means.append(check_y(y, yo, g.G, (), f"{name} step {t} vs oracle"))
        y_means_ref.append(check_y(y, s["y"], g.G, flipped, f"{name} step {t} vs reference"))
    assert np.mean(y_means) <= 1e-3 and np.mean(y_means_ref) <= 1e-3, (np.mean(y_means), np.mean(y_means_ref))
    if prob_total:
        assert prob_mismatch_ref <= max(2, 2e-4 * prob_total)
    # final state against the oracle: bit-exact
    for n in ("pos", "mask", "cache_cts", "k_cache", "v_cache", "attn_history_denom", "attn_counter",
              "attn_history_num", "key_norm", "punc_mask", "num_punc"):
        a, b = getattr(cache, n, None), getattr(oc, n, None)
        if torch.is_tensor(a) and torch.is_tensor(b):
            assert torch.equal(a.cpu().reshape(b.shape).to(b.dtype), b), f"{name}: final {n} differs from the oracle"
    fin = g.state("final")
    assert torch.equal(cache.k_cache.cpu(), fin["k_cache"].reshape(cache.k_cache.shape))
    assert torch.equal(cache.mask.cpu(), fin["mask"].reshape(cache.mask.shape).bool())
    assert torch.equal(cache.cache_cts.cpu(), fin["cache_cts"].reshape(-1).to(torch.int32))


@pytest.mark.parametrize("name", NON_HYBRID)
def test_decode_trace_matches_reference_and_oracle(name):
    _run_trace(name)


@pytest.mark.parametrize("name", HYBRID)
def test_hybrid_decode_trace_matches_reference_and_oracle(name):
    _run_trace(name)


SHAPES = [
    # strategy, Hkv, G, D, L, n_filled, steps
    ("heavy_hitter", 8, 4, 128, 1024, 1024, 48),   # BASELINE config 2 shape, cache full
    ("heavy_hitter", 8, 4, 128, 2036, 2000, 60),   # pyramid layer length (ragged tail, 2 chunks, fill then evict)
    ("heavy_hitter", 4, 7, 128, 556, 556, 40),     # Qwen2-7B grouping, ragged length
    ("heavy_hitter", 1, 4, 128, 1024, 1024, 24),   # one local head (tp.py world size 8)
    ("heavy_hitter", 2, 8, 64, 300, 120, 40),      # D=64, G=8
    ("heavy_hitter:w16", 8, 4, 128, 1024, 1000, 60),  # bounded history window (bf16 ring), fill then evict
    ("heavy_hitter:w4", 2, 2, 64, 200, 60, 80),
    ("heavy_hitter:w8t", 4, 4, 128, 256, 200, 120),    # thresholded (bool) history; oracle-pinned only (the reference raises)
    ("l2", 8, 4, 128, 512, 512, 48),               # BASELINE config 3 shape
    ("l2", 2, 6, 128, 200, 64, 40),
    ("recent_global", 8, 4, 128, 64, 64, 40),      # BASELINE config 1 shape
    ("recent_global", 8, 4, 128, 4096, 4090, 24),
    ("full", 8, 4, 128, 8192, 5000, 12),           # full cache: 8 chunks per CTA through a 4-stage ring
    ("keep_it_odd", 2, 4, 128, 128, 128, 30),
]


@pytest.mark.parametrize("strategy,Hkv,G,D,L,n_filled,steps", SHAPES)
def test_decode_matches_oracle_on_seeded_state(strategy, Hkv, G, D, L, n_filled, steps):
    _seeded_state_run(strategy, Hkv, G, D, L, n_filled, steps)


def _seeded_state_run(strategy, Hkv, G, D, L, n_filled, steps, **step_kwargs):
    from device_utils import make_device_cache, oracle_like, synthetic_fill
    strategy, _, opt = strategy.partition(":")
    extra = dict(W=int(opt.strip("wt")), thresholding=opt.endswith("t")) if opt else {}
    cache = make_device_cache(strategy, Hkv, D, L, **extra)
    next_pos = synthetic_fill(cache, n_filled, seed=L + Hkv, first_pos=0)
    cache._q_per_kv = G
    oc = oracle_like(cache, G)
    gen = torch.Generator().manual_seed(1234 + L)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    y_means = []
    for t in range(steps):
        q = (torch.randn(Hkv * G, D, generator=gen) * 1.5).to(torch.bfloat16)
        k = (torch.randn(Hkv, D, generator=gen) * 1.2).to(torch.bfloat16)
        v = torch.randn(Hkv, D, generator=gen).to(torch.bfloat16)
        pos = next_pos + t
        ip.fill_(pos)
        fi = cache._external_fill_idx(ip).cpu() if cache.policy == 0 else None
        y = cache.decode_step(q.view(1, Hkv * G, 1, D).cuda(), k.view(1, Hkv, 1, D).cuda(),
                              v.view(1, Hkv, 1, D).cuda(), ip, **step_kwargs)[0, :, 0].cpu()
        yo, idx_o, probs_o = oc.decode_step(q, k, v, pos, fill_idx=fi)
        if cache.policy != 0:
            assert torch.equal(cache._fill_idx.cpu(), idx_o), f"eviction index differs at step {t}"
        assert torch.equal(cache.pos[0].cpu(), oc.pos), f"pos differs at step {t}"
        if probs_o is not None:
            assert np.array_equal(bits(cache._out["probs"]), bits(probs_o)), f"pooled probabilities differ at step {t}"
        y_means.append(check_y(y, yo, G, (), f"step {t}"))
    assert np.mean(y_means) <= 1e-3, np.mean(y_means)
    for n in ("mask", "cache_cts", "k_cache", "v_cache", "attn_history_denom", "attn_history_num", "key_norm"):
        a, b = getattr(cache, n, None), getattr(oc, n, None)
        if torch.is_tensor(a) and torch.is_tensor(b):
            assert torch.equal(a.cpu().reshape(b.shape).to(b.dtype), b), f"final {n} differs"
    if hasattr(cache, "attn_history_sum"):  # the kernel's running ring sums stay exact
        assert torch.equal(cache.attn_history_sum.cpu(), cache.attn_history_num.double().sum(-1).cpu())


@pytest.mark.parametrize("W", [1, 6])
def test_three_call_surface_equals_fused_step(W):
    """update_kv -> attend -> update_state (three launches, phase masks) == decode_step (one launch)."""
    from device_utils import make_device_cache, synthetic_fill
    Hkv, G, D, L = 4, 4, 128, 256
    a = make_device_cache("heavy_hitter", Hkv, D, L, W=W)
    b = make_device_cache("heavy_hitter", Hkv, D, L, W=W)
    for c in (a, b):
        synthetic_fill(c, L, seed=7)
        c._q_per_kv = G
    gen = torch.Generator().manual_seed(5)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(20):
        q = (torch.randn(1, Hkv * G, 1, D, generator=gen) * 1.5).to(torch.bfloat16).cuda()
        k = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        v = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        ip.fill_(L + t)
        y_fused = a.decode_step(q, k, v, ip).clone()
        b.update_kv(ip, k, v, False)
        y3, probs = b.attend(q, ip)
        b.update_state(ip, k, v, False, probs.clone())
        assert torch.equal(y_fused, y3)
        for n in ("pos", "mask", "k_cache", "v_cache", "attn_history_num", "attn_history_denom", "cache_cts"):
            assert torch.equal(getattr(a, n), getattr(b, n)), f"{n} differs at step {t}"


def test_full_size_properties():
    """BASELINE config-2 size, no oracle: invariants of the step that hold at any size."""
    from device_utils import make_device_cache, synthetic_fill
    Hkv, G, D, L = 8, 4, 128, 1024
    c = make_device_cache("heavy_hitter", Hkv, D, L)
    synthetic_fill(c, L, seed=3)
    gen = torch.Generator().manual_seed(9)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    den0 = c.attn_history_denom.clone()
    for t in range(64):
        q = (torch.randn(1, Hkv * G, 1, D, generator=gen) * 1.5).to(torch.bfloat16).cuda()
        k = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        v = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        ip.fill_(L + t)
        y = c.decode_step(q, k, v, ip)
        idx = c._fill_idx.clone()
        # the new token sits in the chosen slot of every head, and only there
        assert torch.equal(c.pos[0].gather(1, idx.view(-1, 1)).view(-1), ip.expand(Hkv))
        assert torch.equal(c.k_cache[0].gather(1, idx.view(-1, 1, 1).expand(-1, 1, D))[:, 0], k[0, :, 0])
        # globals (positions < 4) and the 10 most recent positions are never evicted
        assert (c.pos[0] < 4).sum() == 4 * Hkv
        assert ((c.pos[0] >= L + t - 10) & (c.pos[0] <= L + t)).sum() == 11 * Hkv
        # pooled probabilities of a head sum to ~1 (bf16 rounding of 1024 terms)
        s = c._out["probs"].float().sum(-1)
        assert torch.all((s - 1).abs() < 2e-2)
        # y is a convex combination of cached values
        vmax = c.v_cache[0].float().abs().amax(dim=(1, 2)).repeat_interleave(G)
        assert torch.all(y[0, :, 0].float().abs().amax(-1) <= vmax * 1.01)
        assert y.isfinite().all()
    # every slot's denominator advanced by one per step unless it was reset by an eviction
    assert torch.all(c.attn_history_denom <= den0 + 64)
    assert int(c.cache_cts) == L
    assert int(c.attn_counter) == 64
