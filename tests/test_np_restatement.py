"""The plain-numpy restatements (oracle/np_restatement.py: explicit bf16 roundings, no torch arithmetic) against the
vectors recorded from the unmodified reference: quantisation codes / statistics / values bit for bit; prefill attention
probabilities and outputs within the bar of the torch-op restatement."""
import numpy as np
import pytest
import torch

from golden_utils import GOLDEN_DIR, Golden, bits
from oracle import np_restatement as npr
from prefill_utils import PREFILL_FIXTURES, PrefillGolden, assert_close_bf16, check_prefill_y

BITS_FIXTURES = sorted(p.stem for p in GOLDEN_DIR.glob("*_bits*.npz"))


def _f32(t):
    return t.float().numpy()


def test_bf16_round_matches_torch():
    x = np.concatenate([np.random.default_rng(0).standard_normal(100000).astype(np.float32) * 37,
                        np.array([0.0, -0.0, 1.00390625, 1.005859375, 3.3895314e38, 1e-40, 255.5, 256.5], dtype=np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(npr.bf16_round(x), want)


@pytest.mark.parametrize("name", BITS_FIXTURES)
def test_numpy_quantisation_matches_reference(name):
    g = Golden(name)
    n_bit = g.meta["cache_bits"]
    for prefix in ("init", "final"):
        st = g.state(prefix)
        for kv in ("k", "v"):
            ref_codes = st[f"{kv}_cache"].to(torch.int32)
            if n_bit == 8:
                ref_codes = (ref_codes & 0xFF).reshape(g.Hkv, g.L, g.D)
            else:
                shifts = torch.arange(0, 8, n_bit, dtype=torch.int32)
                ref_codes = ((ref_codes.view(-1, 1) >> shifts) & (2 ** n_bit - 1)).reshape(g.Hkv, g.L, g.D)
            dq = npr.dequantize(ref_codes.numpy(), _f32(st[f"{kv}_scales"].reshape(-1)), _f32(st[f"{kv}_zero_points"].reshape(-1)), n_bit)
            assert np.array_equal(dq, _f32(st[f"{kv}_cache_dq"].reshape(g.Hkv, g.L, g.D))), f"{name} {prefix} {kv}: dequantize"
    # one full step of the reference: final state = quantize(cache after the last update); re-quantising the reference's
    # de-quantised final cache with numpy must land on codes within the reference's own drift (values bit-equal after
    # one more round trip through both implementations)
    st = g.state("final")
    for kv in ("k", "v"):
        x = _f32(st[f"{kv}_cache_dq"].reshape(g.Hkv, g.L, g.D))
        codes, scales, zeros = npr.quantize(x, n_bit)
        from oracle import quant
        c2, s2, z2 = quant.quantize(st[f"{kv}_cache_dq"].reshape(g.Hkv, g.L, g.D), n_bit)
        assert np.array_equal(codes, c2.numpy()) and np.array_equal(scales, _f32(s2)) and np.array_equal(zeros, _f32(z2))


@pytest.mark.parametrize("name", [n for n in PREFILL_FIXTURES if "math" not in n])
def test_numpy_prefill_matches_reference(name):
    g = PrefillGolden(name)
    y, pooled = npr.prefill_full_map(_f32(g.t["q"]), _f32(g.t["k"]), _f32(g.t["v"]))
    check_prefill_y(torch.from_numpy(y).reshape(-1, g.D), g.t["y"].reshape(-1, g.D), name)
    rows = g.t["rows_idx"].long().numpy()
    assert_close_bf16(torch.from_numpy(pooled[:, rows, :]), g.t["rows"], f"{name} pooled rows", max_ulp=1.0, min_equal=0.995,
                      floor=2.0 ** -20)
    col = npr.bf16_round(pooled.sum(axis=1, dtype=np.float32))
    assert_close_bf16(torch.from_numpy(col), g.t["col_sum"], f"{name} col_sum", max_ulp=1.0, min_equal=0.97, floor=2.0 ** -14)
