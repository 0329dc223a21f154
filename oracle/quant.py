"""CPU restatement of the reference's KV-cache quantisation round trip.  TEST INFRASTRUCTURE ONLY.

The reference keeps a quantised cache between steps and, in every `update_kv`, de-quantises the WHOLE cache, updates
it, hands the un-quantised tensors to attention and re-quantises everything with fresh statistics
(cache.py:283-338).  Quantisation is asymmetric min/max per SLOT, the group being that slot across all heads and
channels (`quantize_tensor(x, axis=2)` on [1, H, L, D], quantization_utils.py:4-46); every arithmetic step is a bf16
tensor op (fp32 inside, one rounding per op).  Between two steps the cache the next step sees is therefore
`requantize(cache)` = dequantize(quantize(cache)), which is what this module computes.
Pinned on traces recorded from the unmodified reference (tests/golden/*_bits*.npz, tests/test_quant_cache.py).
"""
from __future__ import annotations

import torch


def quantize(x: torch.Tensor, n_bit: int):
    """x [H, L, D] bf16 -> (codes int32 [H, L, D] in [0, 2^n - 1], scales bf16 [L], zero_points bf16 [L])."""
    H, L, D = x.shape
    xt = x.permute(1, 0, 2).reshape(L, H * D)                    # one row per slot: the quantisation group
    mn, mx = torch.aminmax(xt, dim=1)
    max_int = 2 ** n_bit - 1
    scales = (mx - mn).clamp(min=1e-6) / max_int                 # bf16 ops, one rounding each
    zeros = mn + scales * (2 ** (n_bit - 1))
    q = xt.sub(mn.view(-1, 1)).div(scales.view(-1, 1)).round().clamp_(0, max_int)
    codes = q.to(torch.int32).view(L, H, D).permute(1, 0, 2).contiguous()
    return codes, scales, zeros


def dequantize(codes: torch.Tensor, scales: torch.Tensor, zeros: torch.Tensor, n_bit: int) -> torch.Tensor:
    """codes int32 [H, L, D] -> bf16 [H, L, D]: ((code - 2^(n-1)) as bf16) * scale + zero, bf16 ops."""
    c = (codes - 2 ** (n_bit - 1)).to(scales.dtype)
    return c.mul(scales.view(1, -1, 1)).add(zeros.view(1, -1, 1))


def requantize(x: torch.Tensor, n_bit: int) -> torch.Tensor:
    return dequantize(*quantize(x, n_bit), n_bit)


def pack(codes: torch.Tensor, n_bit: int) -> torch.Tensor:
    """Codes -> bytes the way the reference packs them (quantization_utils.py:49-79): consecutive elements of the
    flattened [H, L, D] tensor, first element in the low bits."""
    if n_bit == 8:
        return codes.to(torch.uint8).flatten()
    per = 8 // n_bit
    c = codes.flatten().view(-1, per).to(torch.int32)
    shifts = torch.arange(0, 8, n_bit, dtype=torch.int32)
    return (c << shifts).sum(dim=1).to(torch.uint8)
