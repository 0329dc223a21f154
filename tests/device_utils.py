"""Helpers for the GPU parity tests: build product caches / oracle caches from the same spec."""
from __future__ import annotations

import torch

from cold_compress_b200.cache import get_cache_constructor
from golden_utils import Golden
from oracle.oracle import ATTN_MATH, OracleCache

STATE_NAMES = ("k_cache", "v_cache", "pos", "mask", "cache_cts", "attn_history_num", "attn_history_denom",
               "attn_counter", "key_norm", "special_mask", "punc_mask", "num_special", "num_punc")


def cache_kwargs(strategy, L, g=4, rw=10, hybrid_strategies=None, token_ids=None, min_recovery_frac=0.9, W=1,
                 thresholding=False, bits=None):
    kw = dict(max_cache_length=L, global_tokens=g, max_seq_length=1 << 20, cache_bits=bits, recent_window=rw,
              history_window_size=W, attn_thresholding=thresholding)
    if strategy == "hybrid":
        kw.update(hybrid_strategies=hybrid_strategies, min_recovery_frac=min_recovery_frac,
                  token_ids=token_ids or {"special": [[0], [1]], "punctuation": [5]})
    return kw


def make_device_cache(strategy, Hkv, D, L, device="cuda", **extra):
    cls, relevant = get_cache_constructor(strategy)
    kw = cache_kwargs(strategy, L, **extra)
    return cls(1, Hkv, D, torch.bfloat16, **{k: kw[k] for k in relevant}).to(device)


def load_state(cache, tensors: dict):
    for name in STATE_NAMES:
        dst = getattr(cache, name, None)
        if dst is None or name not in tensors or not torch.is_tensor(dst):
            continue
        src = tensors[name]
        dst.copy_(src.reshape(dst.shape).to(dst.dtype))
    if hasattr(cache, "resync_history"):
        cache.resync_history()


def device_cache_from_golden(g: Golden, device="cuda"):
    m = g.meta
    cache = make_device_cache(m["strategy"], g.Hkv, g.D, g.L, device, g=m["global_tokens"], rw=m["recent_window"],
                              W=m.get("history_window_size", 1), thresholding=m.get("attn_thresholding", False),
                              bits=m.get("cache_bits"), hybrid_strategies=m.get("hybrid_strategies"),
                              token_ids={"special": m.get("special", [[0], [1]]),
                                         "punctuation": m.get("punctuation", [5])})
    init = dict(g.state("init"))
    if m.get("cache_bits"):   # the reference holds integer codes between steps; the product holds the de-quantised view
        init["k_cache"], init["v_cache"] = init.pop("k_cache_dq"), init.pop("v_cache_dq")
    load_state(cache, init)
    if m["strategy"] == "hybrid":
        cache.set_head_strategies(g.z["init_cache_strategies"])
    return cache


def oracle_from_golden(g: Golden) -> OracleCache:
    m = g.meta
    oc = OracleCache(g.policy, g.Hkv, g.G, g.D, g.L, m["global_tokens"], m["recent_window"],
                     m.get("history_window_size", 1), hybrid_strategies=m.get("hybrid_strategies"),
                     attn_thresholding=m.get("attn_thresholding", False))
    oc.load_state(g.state("init"))
    if g.policy == "hybrid":
        oc.set_head_strategies(g.z["init_cache_strategies"])
        if not g.naive():
            oc.attn_mode = ATTN_MATH
    return oc


def oracle_like(cache, G) -> OracleCache:
    """An oracle cache holding a copy of a product cache's current state."""
    pol = {0: "external", 1: "recent_global", 2: "l2", 3: "heavy_hitter", 4: "hybrid"}[cache.policy]
    oc = OracleCache(pol, cache.n_heads, G, cache.head_dim, cache.max_cache_length, int(cache.global_tokens),
                     int(getattr(cache, "recent_window", 0) or 0), int(getattr(cache, "history_window_size", 1)),
                     hybrid_strategies=getattr(cache, "hybrid_strategies", None),
                     attn_thresholding=bool(getattr(cache, "attn_thresholding", False)))
    oc.load_state({n: getattr(cache, n).detach().cpu() for n in STATE_NAMES
                   if torch.is_tensor(getattr(cache, n, None))})
    if pol == "hybrid":
        oc.set_head_strategies(cache.cache_strategies.cpu())
        if not cache.return_attn():
            oc.attn_mode = ATTN_MATH
    return oc


def synthetic_fill(cache, n_filled, seed=0, first_pos=0):
    """Fill the first n_filled slots of a product cache with random K/V at ascending positions."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    H, L, D = cache.n_heads, cache.max_cache_length, cache.head_dim
    dev = cache.k_cache.device
    k = (torch.randn(1, H, n_filled, D, generator=gen) * 1.2).to(torch.bfloat16)
    v = torch.randn(1, H, n_filled, D, generator=gen).to(torch.bfloat16)
    cache.k_cache[:, :, :n_filled] = k.to(dev)
    cache.v_cache[:, :, :n_filled] = v.to(dev)
    pos = torch.arange(first_pos, first_pos + n_filled, dtype=torch.int32)
    cache.pos[:, :, :n_filled] = pos.to(dev)
    cache.mask[:, :, :, :n_filled] = True
    cache.cache_cts.fill_(n_filled)
    if hasattr(cache, "attn_history_num") and cache.attn_history_num.dtype == torch.float64:
        hist = torch.rand(1, H, n_filled, 1, generator=gen, dtype=torch.float64)
        cache.attn_history_num[:, :, :n_filled] = hist.to(dev)
        cache.attn_history_denom[:, :, :n_filled] = torch.randint(1, 50, (1, H, n_filled), generator=gen,
                                                                  dtype=torch.int32).to(dev)
    if hasattr(cache, "key_norm"):
        cache.update_state(None, None, None, True, None)
    return first_pos + n_filled
