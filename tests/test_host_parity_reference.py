"""Host-side (CC_POLICY_EXTERNAL) slot choice of keep_it_odd / random / full against the reference's own classes on
seeded `pos` states (CPU): the index rule of reference cache.py:366-379 with the `_token_importances` of cache.py:519-524
(random) and 1437-1441 (keep_it_odd), and KVCacheFull's first-unfilled rule (cache.py:500-502).  The GPU parity tests
hand this index to the kernel and to the oracle, so the rule itself is pinned here.  Needs the reference checkout
(/root/reference in the build container, baseline/_ref on the GPU box)."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = next((p for p in (Path("/root/reference"), ROOT / "baseline" / "_ref") if (p / "cache.py").exists()), None)
pytestmark = pytest.mark.skipif(REF is None, reason="reference checkout not present")


def _ref_cache_module():
    sys.path.insert(0, str(REF))
    try:
        saved = {n: sys.modules.pop(n) for n in ("cache", "prompt_compression", "quantization_utils") if n in sys.modules}
        spec = importlib.util.spec_from_file_location("ref_cache_external", REF / "cache.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for n in ("cache", "prompt_compression", "quantization_utils"):
            sys.modules.pop(n, None)
        sys.modules.update(saved)
        return mod
    finally:
        sys.path.pop(0)


def _pair(strategy, L, g, rw):
    from cold_compress_b200.cache import get_cache_constructor
    ref = _ref_cache_module()
    kw = dict(max_cache_length=L, max_seq_length=4096, cache_bits=None, global_tokens=g, recent_window=rw)
    rcls, rrel = ref.get_cache_constructor(strategy)
    ocls, orel = get_cache_constructor(strategy)
    assert sorted(rrel) == sorted(orel)
    return rcls(1, 2, 64, torch.bfloat16, **{k: kw[k] for k in rrel}), ocls(1, 2, 64, torch.bfloat16, **{k: kw[k] for k in orel})


@pytest.mark.parametrize("strategy", ["keep_it_odd", "random"])
@pytest.mark.parametrize("L,g,rw,n_filled", [(64, 4, 10, 64), (100, 0, 1, 100), (128, 8, 20, 90), (33, 4, 5, 33)])
def test_external_index_rule_matches_reference(strategy, L, g, rw, n_filled):
    r, o = _pair(strategy, L, g, rw)
    gen = torch.Generator().manual_seed(L * 7 + g)
    for trial in range(40):
        # a plausible cache state: global tokens first, then a random subset of earlier positions in slot order the
        # evictions have scrambled, unfilled slots (-1) at the end while the cache is still filling
        cur = int(torch.randint(n_filled + 50, 3000, (1,), generator=gen))
        kept = torch.randperm(cur - g, generator=gen)[: n_filled - g] + g
        pos = torch.full((L,), -1, dtype=torch.int32)
        pos[:g] = torch.arange(g, dtype=torch.int32)
        pos[g:n_filled] = kept[torch.randperm(n_filled - g, generator=gen)].int()
        r.pos[:] = pos
        o.pos[:] = pos
        ip = torch.tensor([cur], dtype=torch.int32)
        torch.manual_seed(trial)          # `random` draws torch.rand(L) from the global generator in both classes
        want = r._eviction_idx(ip)
        torch.manual_seed(trial)
        got = o._external_fill_idx(ip)
        assert torch.equal(want.view(-1).long(), got.view(-1).long()), f"trial {trial}: reference {want} ours {got}"


def test_full_cache_appends_at_first_unfilled_slot():
    r, o = _pair("full", 96, 0, 0)
    for n in (0, 1, 50, 95):
        pos = torch.full((96,), -1, dtype=torch.int32)
        pos[:n] = torch.arange(n, dtype=torch.int32)
        r.pos[:] = pos
        o.pos[:] = pos
        ip = torch.tensor([n], dtype=torch.int32)
        want = r._eviction_idx(ip) if hasattr(r, "_eviction_idx") else None
        got = o._external_fill_idx(ip)
        assert int(got) == n
        if want is not None:
            assert int(want.view(-1)[0]) == n


@pytest.mark.parametrize("strategy,bits", [("recent_global", None), ("heavy_hitter", None), ("heavy_hitter", 8), ("recent_global", 4),
                                           ("l2", None), ("full", None), ("keep_it_odd", 2)])
def test_cache_statistics_match_reference(strategy, bits):
    """compute_statistics / compression_ratio / memory_usage (reference cache.py:229-281): same numbers from the same
    counters, with and without --cache_bits."""
    from cold_compress_b200.cache import get_cache_constructor
    ref = _ref_cache_module()
    kw = dict(max_cache_length=96, max_seq_length=4096, cache_bits=bits, global_tokens=4, recent_window=10,
              history_window_size=1, attn_thresholding=False)
    rcls, rrel = ref.get_cache_constructor(strategy)
    ocls, orel = get_cache_constructor(strategy)
    assert sorted(rrel) == sorted(orel)
    r = rcls(1, 4, 64, torch.bfloat16, **{k: kw[k] for k in rrel})
    o = ocls(1, 4, 64, torch.bfloat16, **{k: kw[k] for k in orel})
    for filled, seq_len in ((96, 400), (50, 51), (96, 97)):
        r.cache_cts.fill_(filled)
        o.cache_cts.fill_(filled)
        want, got = r.compute_statistics(seq_len), o.compute_statistics(seq_len)
        assert set(want) == set(got)
        assert got["compression_ratio"] == pytest.approx(want["compression_ratio"], rel=1e-6, abs=1e-9)
        if bits is None:
            assert got["cache_memory_gb"] == pytest.approx(want["cache_memory_gb"], rel=1e-6)
        else:
            # --cache_bits here keeps the packed codes AND the bf16 working copy the decode kernel streams (DESIGN.md
            # section 7: the semantics of the quantised cache, not yet its footprint); the number reported is ours
            assert got["cache_memory_gb"] >= want["cache_memory_gb"]
