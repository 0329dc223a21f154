"""Plain-numpy restatements (explicit bf16 roundings, no torch arithmetic) of the two reference paths whose first
restatement in oracle/ is written with torch CPU ops: the quantisation round trip (oracle/quant.py) and the prefill
attention through the full map (oracle/prefill.py).  TEST INFRASTRUCTURE ONLY.

Every bf16 tensor op of the reference is modelled as: operands widened to fp32, the operation in fp32, ONE rounding
to bf16 (round to nearest even).  Pinned on the reference-recorded vectors in tests/golden/ (tests/test_np_restatement.py)."""
from __future__ import annotations

import math

import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest bf16 (ties to even), returned as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    bias = np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    r = ((u + bias) & np.uint32(0xFFFF0000)).astype(np.uint32)
    out = r.view(np.float32).copy()
    nan = np.isnan(x)
    out[nan] = np.nan
    return out


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << np.uint32(16)).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    return (bf16_round(x).view(np.uint32) >> np.uint32(16)).astype(np.uint16)


# ------------------------------------------------------------------ quantisation (cache.py:283-338, quantization_utils.py:4-46)

def quantize(x: np.ndarray, n_bit: int):
    """x fp32 [H, L, D] holding bf16 values -> (codes int32 [H, L, D], scales fp32 [L], zeros fp32 [L])."""
    H, L, D = x.shape
    xt = np.transpose(x, (1, 0, 2)).reshape(L, H * D)
    mn, mx = xt.min(axis=1), xt.max(axis=1)
    max_int = np.float32(2 ** n_bit - 1)
    floor = bf16_round(np.array([1e-6], dtype=np.float32))[0]
    scales = bf16_round(np.maximum(bf16_round(mx - mn), floor) / max_int)
    zeros = bf16_round(mn + bf16_round(scales * np.float32(2 ** (n_bit - 1))))
    q = bf16_round(bf16_round(xt - mn[:, None]) / scales[:, None])
    q = np.clip(np.rint(q), 0, max_int)          # rint: ties to even, like torch.round
    codes = q.astype(np.int32).reshape(L, H, D).transpose(1, 0, 2)
    return np.ascontiguousarray(codes), scales, zeros


def dequantize(codes: np.ndarray, scales: np.ndarray, zeros: np.ndarray, n_bit: int) -> np.ndarray:
    c = (codes - 2 ** (n_bit - 1)).astype(np.float32)
    return bf16_round(bf16_round(c * scales[None, :, None]) + zeros[None, :, None])


# ------------------------------------------------------------------ prefill attention (attention_utils.py:36-54, model.py:413-418)

def prefill_full_map(q: np.ndarray, k: np.ndarray, v: np.ndarray):
    """q fp32 [Hq, S, D], k / v fp32 [Hkv, S, D] holding bf16 values -> (y [Hq, S, D], pooled [Hkv, S, S]), bf16 values."""
    Hq, S, D = q.shape
    Hkv = k.shape[0]
    G = Hq // Hkv
    scale = np.float32(1.0 / math.sqrt(D))
    y = np.zeros((Hq, S, D), dtype=np.float32)
    pooled = np.zeros((Hkv, S, S), dtype=np.float32)
    causal = np.tril(np.ones((S, S), dtype=bool))
    for h in range(Hkv):
        acc = np.zeros((S, S), dtype=np.float32)
        for g in range(G):
            w = bf16_round(q[h * G + g] @ k[h].T)                 # bf16 matmul: fp32 accumulate, one rounding
            w = bf16_round(w * scale)                             # * scale
            w = np.where(causal, w, -np.inf).astype(np.float32)   # + mask bias
            m = w.max(axis=1, keepdims=True)
            e = np.exp((w - m).astype(np.float32)).astype(np.float32)
            p = bf16_round(e / e.sum(axis=1, keepdims=True, dtype=np.float32))   # softmax: fp32 inside, bf16 out
            y[h * G + g] = bf16_round(p @ v[h])
            acc += p
        pooled[h] = bf16_round(acc / np.float32(G))              # mean over the group: fp32 sum / G, bf16 out
    return y, pooled
