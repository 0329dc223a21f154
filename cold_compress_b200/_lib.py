"""ctypes binding of the C ABI in include/cc_b200.h.

There is NO CPU fallback: if libcc_b200.so is missing or does not export every declared symbol,
loading fails loudly.  On a machine without a GPU the library still loads (so the symbol table
can be checked) but every compute call returns CC_ERR_NO_DEVICE / CC_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

SO_PATH = Path(os.environ.get("CC_B200_LIB", PKG / "libcc_b200.so"))  # override: A/B builds in profiles/

ABI_VERSION = 1
CC_OK = 0
POLICY_EXTERNAL, POLICY_RECENT_GLOBAL, POLICY_L2, POLICY_HEAVY_HITTER, POLICY_HYBRID = range(5)
ATTN_MATH_FP32, ATTN_NAIVE_BF16 = 0, 1
PHASE_EVICT_FILL, PHASE_ATTEND, PHASE_SCORE, PHASE_ALL = 1, 2, 4, 7
FLAG_OVERLAP_PREDECESSOR = 1
FLAG_VIA_TOKEN_KERNEL = 2
HYB_FULL, HYB_WINDOW, HYB_HEAVY_HITTER, HYB_SPECIAL, HYB_PUNC = 1, 2, 4, 8, 16

_STATE_INTS = ("policy", "attn_mode", "n_kv_heads", "q_per_kv", "head_dim", "max_cache_length",
               "global_tokens", "recent_window", "history_window", "pos_heads", "cts_heads", "attn_thresholding")
_STATE_PTRS = ("k_cache", "v_cache", "pos", "mask", "cache_cts", "attn_history_num", "attn_history_denom",
               "attn_counter", "attn_history_sum", "key_norm", "hyb_strategy_bits", "hyb_recent_window", "hyb_heavy_budget",
               "special_mask", "punc_mask", "num_special", "num_punc", "sync_words")


class CcCacheState(C.Structure):
    _fields_ = [(n, C.c_int32) for n in _STATE_INTS] + [(n, C.c_void_p) for n in _STATE_PTRS]


class CcDecodeArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("q", "k_new", "v_new", "input_pos", "y", "fill_idx", "probs")] + [
        ("is_punc", C.c_int32), ("phases", C.c_int32), ("flags", C.c_int32), ("n_prefetch", C.c_int32)] + [
        (f"prefetch{i}_{f}", t) for i in range(3) for f, t in (("ptr", C.c_void_p), ("bytes", C.c_uint64))] + [
        ("is_punc_dev", C.c_void_p)]


class CcTpLink(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("world", "rank", "dim", "n_ops")] + [
        (n, C.c_void_p) for n in ("peer_buf", "peer_flag", "cta_done", "epoch")]


class CcTpArgs(C.Structure):
    _fields_ = [("link", C.POINTER(CcTpLink)), ("send_op", C.c_int32), ("recv_op", C.c_int32), ("recv_out", C.c_void_p)]


class CcPrefillAttnArgs(C.Structure):
    _fields_ = [f for t in "qkvy" for f in ((t, C.c_void_p), (f"{t}_stride_s", C.c_int64), (f"{t}_stride_h", C.c_int64))] + [
        (n, C.c_int32) for n in ("seq_len", "n_kv_heads", "q_per_kv", "head_dim", "attn_mode", "obs_len", "n_windows")] + [
        ("windows", C.c_int32 * 4), ("reserved0", C.c_int32)] + [
        (n, C.c_void_p) for n in ("col_sum", "obs_sum", "band_sum", "workspace")] + [("workspace_bytes", C.c_uint64)]


class CcTokenLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("attention_norm", "wqkv", "bqkv", "wo", "ffn_norm", "w1", "w3", "w2")] + [
        ("attention_norm_eps", C.c_float), ("ffn_norm_eps", C.c_float), ("cache", CcCacheState),
        ("fill_idx", C.c_void_p), ("probs", C.c_void_p)]


class CcTokenModel(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layers", "dim", "n_q_heads", "n_kv_heads", "head_dim", "intermediate", "vocab")] + [
        ("norm_eps", C.c_float), ("layers", C.POINTER(CcTokenLayer))] + [
        (n, C.c_void_p) for n in ("tok_embeddings", "norm", "output", "freqs_cis")] + [
        ("world", C.c_int32), ("rank", C.c_int32), ("scratch_peers", C.POINTER(C.c_uint64)), ("scratch_bytes", C.c_uint64)]


# every symbol include/cc_b200.h declares: name -> (restype, argtypes)
EXPORTS = {
    "cc_abi_version": (C.c_int, []),
    "cc_build_info": (C.c_char_p, []),
    "cc_last_error": (C.c_char_p, []),
    "cc_sync_words": (C.c_int, []),
    "cc_decode_step": (C.c_int, [C.POINTER(CcCacheState), C.POINTER(CcDecodeArgs), C.c_void_p]),
    "cc_decode_cluster_size": (C.c_int, [C.POINTER(CcCacheState)]),
    "cc_debug_set_trace": (None, [C.c_void_p]),
    "cc_key_norms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "cc_prefill_gather": (C.c_int, [C.POINTER(CcCacheState), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "cc_requantize_cache": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7),
    "cc_prefill_attention_workspace": (C.c_uint64, [C.c_int32] * 5),
    "cc_prefill_attention": (C.c_int, [C.POINTER(CcPrefillAttnArgs), C.c_void_p]),
    "cc_token_scratch_bytes": (C.c_uint64, [C.POINTER(CcTokenModel)]),
    "cc_token_plan_create": (C.c_int, [C.POINTER(CcTokenModel), C.POINTER(C.c_void_p)]),
    "cc_token_plan_destroy": (C.c_int, [C.c_void_p]),
    "cc_decode_token": (C.c_int, [C.c_void_p] * 7),
    "cc_token_buffer": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]),
    "cc_token_set_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cc_token_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int32)] * 4),
    "cc_rmsnorm_qkv_rope": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.POINTER(CcTpArgs), C.c_void_p]),
    "cc_gemv_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_int32, C.POINTER(CcTpArgs), C.c_void_p]),
    "cc_rmsnorm_ffn_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int32, C.POINTER(CcTpArgs), C.c_void_p]),
}

_lib = None


class CcError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (python -m cold_compress_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        raise CcError(f"{SO_PATH} is missing: build it with `python -m cold_compress_b200.build` "
                      "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the decode path.")
    lib = C.CDLL(str(SO_PATH))
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype, fn.argtypes = restype, argtypes
    if lib.cc_abi_version() != ABI_VERSION:
        raise CcError(f"ABI mismatch: library {lib.cc_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise CcError(f"{what or 'cc call'} failed ({rc}): {load().cc_last_error().decode()}")
    return rc
