"""Prompt compressors (host tensor code) against the reference's classes on the same seeded inputs (CPU):
keep indices, gathered K/V and the cumulative-attention seed (prompt_compression.py:14-209).  Needs the reference
checkout (build container); the recorded prefill vectors in tests/golden/prefill_*.npz cover SnapKV without it."""
import sys
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "prompt_compression.py").exists(), reason="reference checkout not present")
MODS = ("prompt_compression", "cache", "quantization_utils")


def _ref_module():
    for m in MODS:
        sys.modules.pop(m, None)
    sys.path.insert(0, str(REF))
    try:
        import prompt_compression as ref
    finally:
        sys.path.pop(0)
    return ref


@pytest.mark.parametrize("strategy,head_specific", [("recent_global", False), ("recent_global", True), ("l2", True),
                                                    ("keep_it_odd", False), ("heavy_hitter", True), ("full", False)])
def test_compressor_matches_reference(strategy, head_specific):
    from cold_compress_b200 import prompt_compression as ours
    ref = _ref_module()
    try:
        torch.manual_seed(3)
        H, S, D, L = 4, 200, 64, 48
        k = torch.randn(1, H, S, D).to(torch.bfloat16)
        v = torch.randn(1, H, S, D).to(torch.bfloat16)
        pos = torch.arange(S)
        logits = torch.randn(1, H, S, S) * 2
        attn = torch.softmax(logits.masked_fill(~torch.tril(torch.ones(S, S, dtype=torch.bool)), float("-inf")), -1).to(torch.bfloat16)
        kw = dict(max_cache_length=L, global_tokens=4, recent_window=10)
        a = ours.get_prompt_compressor_constructor(strategy)(head_specific=head_specific, **kw)
        b = ref.get_prompt_compressor_constructor(strategy)(head_specific=head_specific, **kw)
        ka, kb = a(pos, k, v, attn=attn), b(pos, k, v, attn=attn)
        assert torch.equal(ka[0], kb[0]), "kept token indices differ"
        assert torch.equal(ka[1], kb[1]) and torch.equal(ka[2], kb[2]), "gathered K/V differ"
        if kb[3] is None:
            assert ka[3] is None
        elif strategy != "full":
            assert torch.equal(ka[3], kb[3]), "cumulative attention seed differs"
        if strategy not in ("full",):
            assert torch.equal(a.select(pos, k, v, attn=attn), kb[0])
    finally:
        for m in MODS:
            sys.modules.pop(m, None)


def test_random_compressor_keeps_globals_and_recent():
    from cold_compress_b200 import prompt_compression as ours
    torch.manual_seed(0)
    S, L = 120, 32
    c = ours.get_prompt_compressor_constructor("random")(head_specific=False, max_cache_length=L, global_tokens=4, recent_window=10)
    k = torch.randn(1, 2, S, 16)
    keep, kk, vv, _ = c(torch.arange(S), k, k)
    assert keep.numel() == L and torch.equal(keep, keep.sort().values)
    assert set(range(4)) <= set(keep.tolist()) and set(range(S - 10, S)) <= set(keep.tolist())
    assert torch.equal(kk, k[:, :, keep])
