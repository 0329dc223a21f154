"""Quantised KV cache (--cache_bits 2/4/8).  Traces *_bits* were recorded from the unmodified reference.

CPU: oracle/quant.py reproduces the reference's quantisation statistics and round trip bit for bit, and the decode
oracle stepped with that round trip between steps reproduces the reference's eviction decisions and outputs.
GPU: the product cache (fused decode kernel + cc_requantize_cache) does."""
import numpy as np
import pytest
import torch

from golden_utils import GOLDEN_DIR, Golden, bf, bits, check_y
from oracle import quant
from oracle.oracle import OracleCache

BITS_FIXTURES = sorted(p.stem for p in GOLDEN_DIR.glob("*_bits*.npz"))


def _oracle(g):
    m = g.meta
    oc = OracleCache(g.policy, g.Hkv, g.G, g.D, g.L, m["global_tokens"], m["recent_window"], 1)
    init = dict(g.state("init"))
    init["k_cache"], init["v_cache"] = init.pop("k_cache_dq"), init.pop("v_cache_dq")   # what the first step sees
    oc.load_state(init)
    return oc


@pytest.mark.parametrize("name", BITS_FIXTURES)
def test_quantisation_round_trip_matches_reference(name):
    g = Golden(name)
    n_bit = g.meta["cache_bits"]
    for prefix in ("init", "final"):
        st = g.state(prefix)
        for kv in ("k", "v"):
            dq = st[f"{kv}_cache_dq"].reshape(g.Hkv, g.L, g.D)
            # the reference's stored codes and statistics reproduce its de-quantised cache through OUR dequantize
            ref_codes = st[f"{kv}_cache"]
            if n_bit == 8:
                ref_codes = (ref_codes.to(torch.int32) & 0xFF).reshape(g.Hkv, g.L, g.D)
            else:
                per = 8 // n_bit
                shifts = torch.arange(0, 8, n_bit, dtype=torch.int32)
                ref_codes = ((ref_codes.to(torch.int32).view(-1, 1) >> shifts) & (2 ** n_bit - 1)).reshape(g.Hkv, g.L, g.D)
                assert torch.equal(quant.pack(ref_codes, n_bit), st[f"{kv}_cache"].to(torch.uint8).flatten())
            mine = quant.dequantize(ref_codes, st[f"{kv}_scales"], st[f"{kv}_zero_points"], n_bit)
            assert torch.equal(mine, dq), f"{name} {prefix} {kv}: dequantize differs from the reference"


@pytest.mark.parametrize("name", BITS_FIXTURES)
def test_oracle_with_requantisation_reproduces_reference_trace(name):
    g = Golden(name)
    n_bit = g.meta["cache_bits"]
    oc = _oracle(g)
    y_means = []
    for t in range(g.T):
        s = g.step(t)
        y, idx, probs = oc.decode_step(s["q"], s["k_new"], s["v_new"], s["input_pos"])
        want = s["fill_idx"].numpy()
        want = np.broadcast_to(want, (g.Hkv,)) if want.size == 1 else want
        assert np.array_equal(idx.numpy(), want), f"{name}: eviction index differs at step {t}"
        flipped = set()
        if s["probs"] is not None:
            diff = np.abs(bits(probs).astype(np.int32) - bits(s["probs"]).astype(np.int32))
            assert diff.max() <= 2
            flipped = set(np.nonzero(diff.reshape(g.Hkv, -1).any(axis=1))[0].tolist())
        y_means.append(check_y(y, s["y"], g.G, flipped, f"{name} step {t}"))
        oc.k_cache.copy_(quant.requantize(oc.k_cache, n_bit))     # cache.py:338: the whole cache, every step
        oc.v_cache.copy_(quant.requantize(oc.v_cache, n_bit))
    assert np.mean(y_means) <= 1e-3
    fin = g.state("final")
    assert torch.equal(oc.k_cache, fin["k_cache_dq"].reshape(oc.k_cache.shape)), "final de-quantised K differs"
    assert torch.equal(oc.v_cache, fin["v_cache_dq"].reshape(oc.v_cache.shape)), "final de-quantised V differs"


@pytest.mark.gpu
@pytest.mark.parametrize("name", BITS_FIXTURES)
def test_quantised_cache_kernel_reproduces_reference_trace(name):
    """Product path: fused decode kernel + cc_requantize_cache per step, against the reference trace and the oracle."""
    from device_utils import device_cache_from_golden
    g = Golden(name)
    n_bit = g.meta["cache_bits"]
    cache = device_cache_from_golden(g)
    oc = _oracle(g)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    y_means = []
    for t in range(g.T):
        s = g.step(t)
        ip.fill_(s["input_pos"])
        y = cache.decode_step(s["q"].view(1, -1, 1, g.D).cuda(), s["k_new"].view(1, -1, 1, g.D).cuda(),
                              s["v_new"].view(1, -1, 1, g.D).cuda(), ip)[0, :, 0].cpu()
        yo, idx_o, probs_o = oc.decode_step(s["q"], s["k_new"], s["v_new"], s["input_pos"])
        oc.k_cache.copy_(quant.requantize(oc.k_cache, n_bit))
        oc.v_cache.copy_(quant.requantize(oc.v_cache, n_bit))
        want = s["fill_idx"].numpy()
        want = np.broadcast_to(want, (g.Hkv,)) if want.size == 1 else want
        assert np.array_equal(cache._fill_idx.cpu().numpy(), want), f"{name}: eviction index differs at step {t}"
        # the re-quantised cache the next step will see: bit-exact against the oracle's round trip
        assert torch.equal(cache.k_cache[0].cpu(), oc.k_cache), f"{name}: re-quantised K differs at step {t}"
        assert torch.equal(cache.v_cache[0].cpu(), oc.v_cache), f"{name}: re-quantised V differs at step {t}"
        flipped = set()
        if s["probs"] is not None:
            assert np.array_equal(bits(cache._out["probs"]), bits(probs_o))
            diff = np.abs(bits(probs_o).astype(np.int32) - bits(s["probs"]).astype(np.int32))
            flipped = set(np.nonzero(diff.reshape(g.Hkv, -1).any(axis=1))[0].tolist())
        check_y(y, yo, g.G, (), f"{name} step {t} vs oracle")
        y_means.append(check_y(y, s["y"], g.G, flipped, f"{name} step {t} vs reference"))
    assert np.mean(y_means) <= 1e-3
    fin = g.state("final")
    assert torch.equal(cache.k_cache.cpu().reshape(-1), fin["k_cache_dq"].reshape(-1))
    assert torch.equal(cache.v_cache.cpu().reshape(-1), fin["v_cache_dq"].reshape(-1))
    # the codes and statistics the reference holds between steps
    for kv in ("k", "v"):
        codes, scales, zeros = quant.quantize(getattr(oc, f"{kv}_cache"), n_bit)
        # (re-quantising the round-tripped cache is what the NEXT update would store; the buffers hold this step's)
        assert torch.equal(getattr(cache, f"{kv}_scales").cpu(), fin[f"{kv}_scales"].reshape(-1))
        assert torch.equal(getattr(cache, f"{kv}_zero_points").cpu(), fin[f"{kv}_zero_points"].reshape(-1))
        ref_bytes = fin[f"{kv}_cache"].to(torch.uint8).flatten() if n_bit < 8 else (fin[f"{kv}_cache"].to(torch.int32) & 0xFF).to(torch.uint8).flatten()
        assert torch.equal(getattr(cache, f"{kv}_codes").cpu(), ref_bytes), f"{name}: packed {kv} codes differ from the reference's"


@pytest.mark.gpu
def test_quantised_cache_three_call_surface_and_errors():
    from cold_compress_b200 import _lib
    from device_utils import make_device_cache, synthetic_fill
    Hkv, G, D, L = 4, 4, 128, 128
    a = make_device_cache("heavy_hitter", Hkv, D, L, bits=4)
    b = make_device_cache("heavy_hitter", Hkv, D, L, bits=4)
    for c in (a, b):
        synthetic_fill(c, L, seed=7)
        c.requantize()
        c._q_per_kv = G
    gen = torch.Generator().manual_seed(5)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(12):
        q = (torch.randn(1, Hkv * G, 1, D, generator=gen) * 1.5).to(torch.bfloat16).cuda()
        k = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        v = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        ip.fill_(L + t)
        y_fused = a.decode_step(q, k, v, ip).clone()
        b.update_kv(ip, k, v, False)
        y3, probs = b.attend(q, ip)
        b.update_state(ip, k, v, False, probs.clone())
        assert torch.equal(y_fused, y3)
        for n in ("pos", "k_cache", "v_cache", "k_codes", "v_scales", "attn_history_num"):
            assert torch.equal(getattr(a, n), getattr(b, n)), f"{n} differs at step {t}"
    assert a.k_codes.numel() == Hkv * L * D // 2
    with pytest.raises(NotImplementedError):
        make_device_cache("l2", Hkv, D, L, bits=8)
    with pytest.raises(_lib.CcError):
        from cold_compress_b200 import ops
        ops.requantize_cache(a.k_cache, a.v_cache, 3, a.k_codes, a.v_codes, a.k_scales, a.k_zero_points, a.v_scales, a.v_zero_points)
