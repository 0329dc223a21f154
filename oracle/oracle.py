"""ctypes front-end of the CPU oracle (oracle/cc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  It holds one layer's cache state in host tensors named like
the reference's buffers (cache.py:195-227, 574, 668-679) and steps it with the C oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent

POLICY = {"external": 0, "recent_global": 1, "l2": 2, "heavy_hitter": 3, "hybrid": 4}
ATTN_MATH, ATTN_NAIVE = 0, 1
PHASE_EVICT_FILL, PHASE_ATTEND, PHASE_SCORE, PHASE_ALL = 1, 2, 4, 7
HYB_BITS = {"full": 1, "window": 2, "heavy_hitter": 4, "special": 8, "punc": 16}


class CacheState(C.Structure):
    """Mirror of CcCacheState (include/cc_b200.h)."""

    _fields_ = [(n, C.c_int32) for n in (
        "policy", "attn_mode", "n_kv_heads", "q_per_kv", "head_dim", "max_cache_length",
        "global_tokens", "recent_window", "history_window", "pos_heads", "cts_heads", "attn_thresholding")] + [
        (n, C.c_void_p) for n in (
            "k_cache", "v_cache", "pos", "mask", "cache_cts", "attn_history_num",
            "attn_history_denom", "attn_counter", "attn_history_sum", "key_norm", "hyb_strategy_bits",
            "hyb_recent_window", "hyb_heavy_budget", "special_mask", "punc_mask",
            "num_special", "num_punc", "sync_words")]


class DecodeArgs(C.Structure):
    """Mirror of CcDecodeArgs (include/cc_b200.h)."""

    _fields_ = [(n, C.c_void_p) for n in ("q", "k_new", "v_new", "input_pos", "y", "fill_idx", "probs")] + [
        ("is_punc", C.c_int32), ("phases", C.c_int32), ("flags", C.c_int32), ("n_prefetch", C.c_int32)] + [
        (f"prefetch{i}_{f}", t) for i in range(3) for f, t in (("ptr", C.c_void_p), ("bytes", C.c_uint64))] + [
        ("is_punc_dev", C.c_void_p)]


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "cc_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return so


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.cc_oracle_decode_step.argtypes = [C.POINTER(CacheState), C.POINTER(DecodeArgs)]
        _lib.cc_oracle_decode_step.restype = C.c_int
        _lib.cc_oracle_key_norms.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        _lib.cc_oracle_key_norms.restype = C.c_int
        _lib.cc_oracle_last_error.restype = C.c_char_p
        _lib.cc_oracle_expf.argtypes = [C.c_float]
        _lib.cc_oracle_expf.restype = C.c_float
        _lib.cc_oracle_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.cc_oracle_dot.restype = C.c_float
        _lib.cc_oracle_threads.restype = C.c_int
    return _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def strategy_bits(name: str) -> int:
    """CC_HYB_* bits of a hybrid strategy name (substring tests of cache.py:896-950)."""
    if name == "full":
        return HYB_BITS["full"]
    return sum(bit for key, bit in HYB_BITS.items() if key != "full" and key in name)


class OracleCache:
    """One layer's state on the host + the oracle step.  Buffers without the batch dim."""

    def __init__(self, policy: str, n_kv_heads: int, q_per_kv: int, head_dim: int, max_cache_length: int,
                 global_tokens: int = 4, recent_window: int = 10, history_window: int = 1,
                 hybrid_strategies=None, attn_mode: int | None = None, attn_thresholding: bool = False):
        H, L, D = n_kv_heads, max_cache_length, head_dim
        self.policy, self.H, self.G, self.D, self.L = policy, H, q_per_kv, D, L
        self.global_tokens, self.recent_window = global_tokens, recent_window
        head_specific = policy in ("l2", "heavy_hitter", "hybrid")
        self.pos_heads = H if head_specific else 1
        self.cts_heads = H if policy == "hybrid" else 1
        self.W = 400 if policy == "hybrid" else history_window
        self.attn_thresholding = bool(attn_thresholding) and policy == "heavy_hitter"
        self.attn_mode = attn_mode if attn_mode is not None else (
            ATTN_NAIVE if policy in ("heavy_hitter", "hybrid") else ATTN_MATH)
        self.k_cache = torch.zeros(H, L, D, dtype=torch.bfloat16)
        self.v_cache = torch.zeros(H, L, D, dtype=torch.bfloat16)
        self.pos = torch.full((self.pos_heads, L), -1, dtype=torch.int32)
        self.mask = torch.zeros(H, L, dtype=torch.bool)
        self.cache_cts = torch.zeros(self.cts_heads, dtype=torch.int32)
        self.attn_history_num = self.attn_history_denom = self.attn_counter = None
        self.key_norm = None
        self.hyb_bits = self.hyb_rw = self.hyb_hh = None
        self.special_mask = self.punc_mask = self.num_special = self.num_punc = None
        if policy in ("heavy_hitter", "hybrid"):
            self.attn_history_num = (torch.zeros(H, L, dtype=torch.float64) if self.W == 1
                                     else torch.zeros(H, L, self.W, dtype=torch.bool if self.attn_thresholding
                                                      else torch.bfloat16))
            self.attn_history_denom = torch.zeros(H, L, dtype=torch.int32)
            self.attn_counter = torch.zeros(1, dtype=torch.int64)
        if policy == "l2":
            self.key_norm = torch.zeros(H, L, dtype=torch.bfloat16)
        if policy == "hybrid":
            self.hybrid_strategies = hybrid_strategies
            self.hyb_bits = torch.zeros(H, dtype=torch.int32)
            self.hyb_rw = torch.zeros(H, dtype=torch.int32)
            self.hyb_hh = torch.zeros(H, dtype=torch.int32)
            self.special_mask = torch.zeros(H, L, dtype=torch.bool)
            self.punc_mask = torch.zeros(H, L, dtype=torch.bool)
            self.num_special = torch.zeros(1, dtype=torch.int32)
            self.num_punc = torch.zeros(1, dtype=torch.int32)

    def set_head_strategies(self, strategy_ids):
        """strategy_ids[h] indexes hybrid_strategies (cache.py:983: self.cache_strategies)."""
        for h, sid in enumerate(strategy_ids):
            s = self.hybrid_strategies[int(sid)]
            self.hyb_bits[h] = strategy_bits(s["strategy"])
            self.hyb_rw[h] = round(s.get("recent_window", 0) * self.L)
            self.hyb_hh[h] = round(s.get("heavy_hitter_frac", 0) * self.L)

    def state(self) -> CacheState:
        st = CacheState()
        st.policy, st.attn_mode = POLICY[self.policy], self.attn_mode
        st.n_kv_heads, st.q_per_kv, st.head_dim, st.max_cache_length = self.H, self.G, self.D, self.L
        st.global_tokens, st.recent_window, st.history_window = self.global_tokens, self.recent_window, self.W
        st.pos_heads, st.cts_heads = self.pos_heads, self.cts_heads
        st.attn_thresholding = int(self.attn_thresholding)
        for name, t in (("k_cache", self.k_cache), ("v_cache", self.v_cache), ("pos", self.pos),
                        ("mask", self.mask), ("cache_cts", self.cache_cts),
                        ("attn_history_num", self.attn_history_num),
                        ("attn_history_denom", self.attn_history_denom), ("attn_counter", self.attn_counter),
                        ("key_norm", self.key_norm), ("hyb_strategy_bits", self.hyb_bits),
                        ("hyb_recent_window", self.hyb_rw), ("hyb_heavy_budget", self.hyb_hh),
                        ("special_mask", self.special_mask), ("punc_mask", self.punc_mask),
                        ("num_special", self.num_special), ("num_punc", self.num_punc)):
            setattr(st, name, _ptr(t))
        return st

    def decode_step(self, q, k_new, v_new, input_pos: int, fill_idx=None, is_punc: bool = False,
                    phases: int = PHASE_ALL, probs_in=None):
        """q [H*G, D], k_new/v_new [H, D] bf16 host tensors.  Returns (y, fill_idx, probs)."""
        H, G, D, L = self.H, self.G, self.D, self.L
        q = q.contiguous().view(H * G, D)
        k_new, v_new = k_new.contiguous().view(H, D), v_new.contiguous().view(H, D)
        ip = torch.tensor([input_pos], dtype=torch.int32)
        y = torch.zeros(H * G, D, dtype=torch.bfloat16)
        if self.policy == "external":
            idx = fill_idx.to(torch.int64).contiguous().clone()
        else:
            idx = torch.full((H,), -1, dtype=torch.int64)
        probs = None
        if self.attn_mode == ATTN_NAIVE:
            probs = probs_in.contiguous().clone() if probs_in is not None else torch.zeros(H, L, dtype=torch.bfloat16)
        a = DecodeArgs()
        a.q, a.k_new, a.v_new, a.input_pos, a.y = _ptr(q), _ptr(k_new), _ptr(v_new), _ptr(ip), _ptr(y)
        a.fill_idx, a.probs = _ptr(idx), _ptr(probs)
        a.is_punc, a.phases = int(is_punc), phases
        st = self.state()
        rc = lib().cc_oracle_decode_step(C.byref(st), C.byref(a))
        if rc != 0:
            raise RuntimeError(f"oracle: {lib().cc_oracle_last_error().decode()} ({rc})")
        return y, idx, probs

    def seed_key_norms(self):
        rc = lib().cc_oracle_key_norms(_ptr(self.k_cache), _ptr(self.key_norm), self.H * self.L, self.D)
        assert rc == 0

    STATE_NAMES = ("k_cache", "v_cache", "pos", "mask", "cache_cts", "attn_history_num",
                   "attn_history_denom", "attn_counter", "key_norm", "special_mask", "punc_mask",
                   "num_special", "num_punc")

    def load_state(self, tensors: dict):
        """Copy reference-shaped tensors (with or without the leading batch dim) into this state."""
        for name in self.STATE_NAMES:
            dst = getattr(self, name)
            if dst is None or name not in tensors:
                continue
            src = tensors[name]
            src = src if torch.is_tensor(src) else torch.as_tensor(src)
            dst.copy_(src.reshape(dst.shape).to(dst.dtype))


def threads() -> int:
    return lib().cc_oracle_threads()
