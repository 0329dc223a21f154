"""PyTorch custom ops over the C ABI (include/cc_b200.h).

`torch.ops.cold_compress_b200.decode_step` is the fused per-layer decode step; it mutates the
cache buffers in place and is registered with a schema that says so, so that functionalisation
and CUDA-graph capture see the writes.  Ops exist for CUDA tensors only: calling them with CPU
tensors raises (there is no CPU path in the product).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_LIB_NAME = "cold_compress_b200"
_library = torch.library.Library(_LIB_NAME, "DEF")

# order of the tensors in `state` (empty tensor = NULL); must match _lib._STATE_PTRS
STATE_ORDER = _lib._STATE_PTRS
META_ORDER = _lib._STATE_INTS

_library.define(
    "decode_step(Tensor q, Tensor k_new, Tensor v_new, Tensor input_pos, Tensor(a!)[] state, int[] meta, "
    "Tensor(b!) y, Tensor(c!) fill_idx, Tensor(d!) probs, int phases, bool is_punc, int flags, Tensor[] warm, Tensor? is_punc_dev=None) -> ()")
_library.define("key_norms(Tensor k, Tensor(a!) key_norm) -> ()")
_library.define(
    "prefill_gather(Tensor(a!)[] state, int[] meta, Tensor k_prompt, Tensor v_prompt, Tensor keep_idx) -> ()")
_library.define("requantize_cache(Tensor(a!) k_cache, Tensor(b!) v_cache, int n_bit, Tensor(c!) k_codes, Tensor(d!) v_codes, "
                "Tensor(e!) k_scales, Tensor(f!) k_zero_points, Tensor(g!) v_scales, Tensor(h!) v_zero_points) -> ()")
_library.define(
    "prefill_attention(Tensor q, Tensor k, Tensor v, Tensor(a!) y, int attn_mode, int obs_len, int[] windows, "
    "Tensor(b!)? col_sum, Tensor(c!)? obs_sum, Tensor(d!)? band_sum, Tensor(e!)? workspace) -> ()")
_library.define(
    "rmsnorm_qkv_rope(Tensor x, Tensor norm_w, float eps, Tensor wqkv, Tensor? bias, Tensor freqs_cis, "
    "Tensor input_pos, Tensor(a!) q, Tensor(b!) k, Tensor(c!) v, int n_q_heads, int n_kv_heads, int head_dim, "
    "Tensor[] tp_link, int[] tp_meta, Tensor(d!)? recv_out) -> ()")
_library.define("gemv_residual(Tensor x, Tensor? norm_w, float eps, Tensor w, Tensor? residual, Tensor(a!) out, "
                "Tensor[] tp_link, int[] tp_meta, Tensor(b!)? recv_out) -> ()")
_library.define("rmsnorm_ffn_gate(Tensor x, Tensor norm_w, float eps, Tensor w1, Tensor w3, Tensor(a!) out, "
                "Tensor[] tp_link, int[] tp_meta, Tensor(b!)? recv_out) -> ()")


def _ptr(t: torch.Tensor):
    return t.data_ptr() if t.numel() else None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_state_cache: dict = {}


def _c_state(state, meta) -> _lib.CcCacheState:
    key = (tuple(t.data_ptr() for t in state), tuple(meta))
    st = _state_cache.get(key)
    if st is None:
        st = _lib.CcCacheState()
        for name, v in zip(META_ORDER, meta):
            setattr(st, name, int(v))
        for name, t in zip(STATE_ORDER, state):
            setattr(st, name, _ptr(t))
        if len(_state_cache) > 4096:
            _state_cache.clear()
        _state_cache[key] = st
    return st


def _require_cuda(*tensors):
    for t in tensors:
        if t.numel() and not t.is_cuda:
            raise _lib.CcError("cold_compress_b200 ops run on CUDA tensors only (no CPU fallback)")
        if t.numel() and not t.is_contiguous():
            raise _lib.CcError("cold_compress_b200 ops need contiguous tensors")


def _decode_step(q, k_new, v_new, input_pos, state, meta, y, fill_idx, probs, phases, is_punc, flags=0, warm=(), is_punc_dev=None):
    _require_cuda(q, k_new, v_new, input_pos, y, fill_idx, probs, *state)
    lib = _lib.load()
    st = _c_state(state, meta)
    a = _lib.CcDecodeArgs()
    a.q, a.k_new, a.v_new, a.input_pos = _ptr(q), _ptr(k_new), _ptr(v_new), _ptr(input_pos)
    a.y, a.fill_idx, a.probs = _ptr(y), _ptr(fill_idx), _ptr(probs)
    a.is_punc, a.phases, a.flags = int(is_punc), int(phases), int(flags)
    a.is_punc_dev = None if is_punc_dev is None else is_punc_dev.data_ptr()
    a.n_prefetch = min(3, len(warm))
    for i, t in enumerate(warm[:3]):  # L2 warm-up hints: contiguous read-only tensors (views allowed)
        setattr(a, f"prefetch{i}_ptr", t.data_ptr())
        setattr(a, f"prefetch{i}_bytes", (t.numel() * t.element_size()) // 16 * 16)
    _lib.check(lib.cc_decode_step(C.byref(st), C.byref(a), _stream()), "cc_decode_step")


def _key_norms(k, key_norm):
    _require_cuda(k, key_norm)
    lib = _lib.load()
    head_dim = k.shape[-1]
    _lib.check(lib.cc_key_norms(_ptr(k), _ptr(key_norm), k.numel() // head_dim, head_dim, _stream()), "cc_key_norms")


def _prefill_gather(state, meta, k_prompt, v_prompt, keep_idx):
    _require_cuda(k_prompt, v_prompt, keep_idx, *state)
    lib = _lib.load()
    st = _c_state(state, meta)
    keep_heads = 1 if keep_idx.dim() == 1 else keep_idx.shape[0]
    n_keep = keep_idx.shape[-1]
    prompt_len = k_prompt.shape[-2]
    _lib.check(lib.cc_prefill_gather(C.byref(st), _ptr(k_prompt), _ptr(v_prompt), _ptr(keep_idx), keep_heads,
                                     n_keep, prompt_len, _stream()), "cc_prefill_gather")


def _requantize_cache(k_cache, v_cache, n_bit, k_codes, v_codes, k_scales, k_zero_points, v_scales, v_zero_points):
    """k_cache / v_cache bf16 [1, H, L, D] (or [H, L, D]) contiguous, re-quantised in place."""
    _require_cuda(k_cache, v_cache, k_codes, v_codes, k_scales, k_zero_points, v_scales, v_zero_points)
    H, L, D = k_cache.shape[-3:]
    _lib.check(_lib.load().cc_requantize_cache(k_cache.data_ptr(), v_cache.data_ptr(), H, L, D, int(n_bit), _ptr(k_codes),
                                               _ptr(v_codes), _ptr(k_scales), _ptr(k_zero_points), _ptr(v_scales),
                                               _ptr(v_zero_points), _stream()), "cc_requantize_cache")


def prefill_attention_workspace(seq_len, n_kv_heads, q_per_kv, obs_len, n_windows) -> int:
    return int(_lib.load().cc_prefill_attention_workspace(seq_len, n_kv_heads, q_per_kv, obs_len, n_windows))


def _prefill_attention(q, k, v, y, attn_mode, obs_len, windows, col_sum, obs_sum, band_sum, workspace):
    """q / y [Hq, S, D] and k / v [Hkv, S, D] VIEWS (any row / head strides, D contiguous): cc_prefill_attention."""
    for t in (q, k, v, y, col_sum, obs_sum, band_sum, workspace):
        if t is not None and t.numel() and not t.is_cuda:
            raise _lib.CcError("cold_compress_b200 ops run on CUDA tensors only (no CPU fallback)")
    for t in (q, k, v, y):
        if t.dim() != 3 or t.stride(2) != 1 or t.dtype != torch.bfloat16:
            raise _lib.CcError("prefill_attention: expected bf16 [heads, S, D] views with a contiguous last dimension")
    a = _lib.CcPrefillAttnArgs()
    for name, t in (("q", q), ("k", k), ("v", v), ("y", y)):
        setattr(a, name, t.data_ptr())
        setattr(a, name + "_stride_h", t.stride(0))
        setattr(a, name + "_stride_s", t.stride(1))
    a.seq_len, a.n_kv_heads, a.q_per_kv, a.head_dim = q.shape[1], k.shape[0], q.shape[0] // k.shape[0], q.shape[2]
    a.attn_mode, a.obs_len, a.n_windows = int(attn_mode), int(obs_len), len(windows)
    for i, w in enumerate(windows):
        a.windows[i] = int(w)
    a.col_sum, a.obs_sum, a.band_sum = _optr(col_sum), _optr(obs_sum), _optr(band_sum)
    a.workspace = _optr(workspace)
    a.workspace_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    _lib.check(_lib.load().cc_prefill_attention(C.byref(a), _stream()), "cc_prefill_attention")


def _optr(t):
    return None if t is None else t.data_ptr()


def _tp_args(tp_link, tp_meta, recv_out):
    """tp_link = [peer_buf, peer_flag, cta_done, epoch] device tensors (empty list = no exchange),
    tp_meta = [world, rank, dim, n_ops, send_op, recv_op]."""
    if not tp_link:
        return None, None
    link = _lib.CcTpLink()
    link.world, link.rank, link.dim, link.n_ops = (int(v) for v in tp_meta[:4])
    link.peer_buf, link.peer_flag, link.cta_done, link.epoch = (t.data_ptr() for t in tp_link)
    args = _lib.CcTpArgs()
    args.link = C.pointer(link)
    args.send_op, args.recv_op = int(tp_meta[4]), int(tp_meta[5])
    args.recv_out = None if recv_out is None else recv_out.data_ptr()
    return args, link  # keep `link` alive until the call returns


def _rmsnorm_qkv_rope(x, norm_w, eps, wqkv, bias, freqs_cis, input_pos, q, k, v, n_q_heads, n_kv_heads, head_dim,
                      tp_link=(), tp_meta=(), recv_out=None):
    _require_cuda(x, norm_w, wqkv, freqs_cis, input_pos, q, k, v)
    tp, _keep = _tp_args(tp_link, tp_meta, recv_out)
    _lib.check(_lib.load().cc_rmsnorm_qkv_rope(
        x.data_ptr(), norm_w.data_ptr(), eps, wqkv.data_ptr(), _optr(bias), freqs_cis.data_ptr(), input_pos.data_ptr(),
        q.data_ptr(), k.data_ptr(), v.data_ptr(), wqkv.shape[1], n_q_heads, n_kv_heads, head_dim,
        None if tp is None else C.byref(tp), _stream()), "cc_rmsnorm_qkv_rope")


def _gemv_residual(x, norm_w, eps, w, residual, out, tp_link=(), tp_meta=(), recv_out=None):
    _require_cuda(x, w, out)
    tp, _keep = _tp_args(tp_link, tp_meta, recv_out)
    _lib.check(_lib.load().cc_gemv_residual(x.data_ptr(), _optr(norm_w), eps, w.data_ptr(), _optr(residual),
                                            out.data_ptr(), w.shape[1], w.shape[0],
                                            None if tp is None else C.byref(tp), _stream()), "cc_gemv_residual")


def _rmsnorm_ffn_gate(x, norm_w, eps, w1, w3, out, tp_link=(), tp_meta=(), recv_out=None):
    _require_cuda(x, norm_w, w1, w3, out)
    tp, _keep = _tp_args(tp_link, tp_meta, recv_out)
    _lib.check(_lib.load().cc_rmsnorm_ffn_gate(x.data_ptr(), norm_w.data_ptr(), eps, w1.data_ptr(), w3.data_ptr(),
                                               out.data_ptr(), w1.shape[1], w1.shape[0],
                                               None if tp is None else C.byref(tp), _stream()), "cc_rmsnorm_ffn_gate")


_library.impl("decode_step", _decode_step, "CUDA")
_library.impl("rmsnorm_qkv_rope", _rmsnorm_qkv_rope, "CUDA")
_library.impl("gemv_residual", _gemv_residual, "CUDA")
_library.impl("rmsnorm_ffn_gate", _rmsnorm_ffn_gate, "CUDA")
_library.impl("key_norms", _key_norms, "CUDA")
_library.impl("prefill_gather", _prefill_gather, "CUDA")
_library.impl("prefill_attention", _prefill_attention, "CUDA")
_library.impl("requantize_cache", _requantize_cache, "CUDA")


def _no_cpu(*_a, **_k):
    raise _lib.CcError("cold_compress_b200 ops have no CPU implementation: the decode path is CUDA (sm_100a) only")


for _name in ("decode_step", "key_norms", "prefill_gather", "prefill_attention", "requantize_cache", "rmsnorm_qkv_rope", "gemv_residual", "rmsnorm_ffn_gate"):
    _library.impl(_name, _no_cpu, "CPU")

decode_step = torch.ops.cold_compress_b200.decode_step
key_norms = torch.ops.cold_compress_b200.key_norms
prefill_gather = torch.ops.cold_compress_b200.prefill_gather
prefill_attention = torch.ops.cold_compress_b200.prefill_attention
requantize_cache = torch.ops.cold_compress_b200.requantize_cache
rmsnorm_qkv_rope = torch.ops.cold_compress_b200.rmsnorm_qkv_rope
gemv_residual = torch.ops.cold_compress_b200.gemv_residual
rmsnorm_ffn_gate = torch.ops.cold_compress_b200.rmsnorm_ffn_gate
