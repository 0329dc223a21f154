"""Host side of the whole-token persistent kernel (csrc/decode_token.cu, `cc_decode_token`).

`TokenPlan(model)` describes a `Transformer` whose caches are set up (weights, per-layer cache buffers, the RoPE
table) to the C ABI once; `plan.decode(token, input_pos)` then runs ONE kernel per decoded token: embedding row ->
32 x [attention_norm + wqkv + RoPE | eviction + overwrite + attention + score update | wo | ffn gate | w2] -> norm ->
output head -> arg-max token and softmax probabilities (reference generation_utils.py:163-178 + 136-142).
Under tensor parallelism (tp.apply_tp) the scratch is symmetric (peer-mapped) memory and the all-reduces of
tp.py:134-160 are stores into every rank's scratch from inside the same kernel.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_SUPPORTED = {_lib.POLICY_RECENT_GLOBAL, _lib.POLICY_L2, _lib.POLICY_HEAVY_HITTER}


def unsupported_reason(model) -> str | None:
    """None if `model` (a Transformer with caches) can be decoded by the whole-token kernel, else why not."""
    cfg = model.config
    att0 = model.layers[0].attention
    if att0.kv_cache is None:
        return "caches are not set up"
    w = model.output.weight
    if not w.is_cuda or w.dtype != torch.bfloat16:
        return "needs bf16 weights on a CUDA device"
    dim = model.output.in_features
    if dim % 256 or (att0.n_head * att0.head_dim) % 256 or model.layers[0].feed_forward.w1.out_features % 256:
        return "dim, n_head * head_dim and the (local) intermediate size must be multiples of 256"
    if att0.head_dim not in (64, 128) or att0.n_head // att0.n_local_heads not in (1, 2, 4, 7, 8):
        return "head_dim must be 64 or 128 and n_head / n_local_heads one of 1, 2, 4, 7, 8"
    if model.output.out_features % 2:
        return "odd vocabulary size"
    for layer in model.layers:
        c = layer.attention.kv_cache
        if type(c).__name__ == "KVCacheAnalysis" or c.policy not in _SUPPORTED:
            return f"cache policy of {type(c).__name__} runs through the per-layer kernel"
        if c.quantize:
            return "--cache_bits runs through the per-layer kernel"
        if c.policy == _lib.POLICY_HEAVY_HITTER and (c.history_window_size != 1 or getattr(c, "attn_thresholding", False)):
            return "heavy_hitter with a history window / thresholding runs through the per-layer kernel"
    return None


class TokenPlan:
    def __init__(self, model, group=None):
        why = unsupported_reason(model)
        if why:
            raise _lib.CcError(f"whole-token kernel: {why}")
        self.lib = _lib.load()
        self.model = model
        dev = model.output.weight.device
        n = len(model.layers)
        att0 = model.layers[0].attention
        G = att0.n_head // att0.n_local_heads
        self._layers = (_lib.CcTokenLayer * n)()
        self._keep = []
        for i, block in enumerate(model.layers):
            att, ffn, cache = block.attention, block.feed_forward, block.attention.kv_cache
            out = cache._outputs(G, dev, cache.k_cache.dtype)
            cache._q_per_kv = G
            ly = self._layers[i]
            ly.attention_norm = block.attention_norm.weight.data_ptr()
            ly.wqkv = att.wqkv.weight.data_ptr()
            ly.bqkv = att.wqkv.bias.data_ptr() if att.wqkv.bias is not None else None
            ly.wo = att.wo.weight.data_ptr()
            ly.ffn_norm = block.ffn_norm.weight.data_ptr()
            ly.w1, ly.w3, ly.w2 = ffn.w1.weight.data_ptr(), ffn.w3.weight.data_ptr(), ffn.w2.weight.data_ptr()
            ly.attention_norm_eps, ly.ffn_norm_eps = block.attention_norm.eps, block.ffn_norm.eps
            st = ly.cache
            for name, v in zip(_lib._STATE_INTS, cache._meta(G)):
                setattr(st, name, int(v))
            for name, t in zip(_lib._STATE_PTRS, cache._state_list()):
                setattr(st, name, t.data_ptr() if t.numel() else None)
            ly.fill_idx = cache._fill_idx.data_ptr()
            ly.probs = out["probs"].data_ptr() if cache.return_attn() else None
        m = _lib.CcTokenModel()
        m.n_layers, m.dim = n, model.output.in_features
        m.n_q_heads, m.n_kv_heads, m.head_dim = att0.n_head, att0.n_local_heads, att0.head_dim
        m.intermediate = model.layers[0].feed_forward.w1.out_features
        m.vocab = model.output.out_features
        m.norm_eps = model.norm.eps
        m.layers = C.cast(self._layers, C.POINTER(_lib.CcTokenLayer))
        m.tok_embeddings = model.tok_embeddings.weight.data_ptr()
        m.norm = model.norm.weight.data_ptr()
        m.output = model.output.weight.data_ptr()
        m.freqs_cis = model.freqs_cis.data_ptr()
        self.world, self.rank = 1, 0
        tp_group = att0.tp_group
        if tp_group is not None:
            import torch.distributed as dist
            group = group or dist.group.WORLD
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        m.world, m.rank = self.world, self.rank
        peers = (C.c_uint64 * max(1, self.world))()
        m.scratch_peers = C.cast(peers, C.POINTER(C.c_uint64))
        nbytes = int(self.lib.cc_token_scratch_bytes(C.byref(m)))
        if nbytes == 0:
            raise _lib.CcError(f"cc_token_scratch_bytes failed: {self.lib.cc_last_error().decode()}")
        if self.world == 1:
            self.scratch = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            peers[0] = self.scratch.data_ptr()
        else:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            self.scratch = symm.empty(nbytes, dtype=torch.uint8, device=dev)
            self.scratch.zero_()
            self._handle = symm.rendezvous(self.scratch, group.group_name)
            torch.cuda.synchronize()
            dist.barrier(group)
            for r, ptr in enumerate(self._handle.buffer_ptrs):
                peers[r] = ptr
        m.scratch_bytes = nbytes
        self._m, self._peers = m, peers
        handle = C.c_void_p()
        _lib.check(self.lib.cc_token_plan_create(C.byref(m), C.byref(handle)), "cc_token_plan_create")
        self.handle = handle
        self.token_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits = torch.empty(m.vocab, dtype=torch.bfloat16, device=dev)
        self.probs = torch.empty(m.vocab, dtype=torch.bfloat16, device=dev)
        self.signature = self.signature_of(model)

    @staticmethod
    def signature_of(model):
        """Changes when a plan built earlier no longer describes the model (caches rebuilt, weights moved)."""
        sig = [model.output.weight.data_ptr(), model.freqs_cis.data_ptr() if model.freqs_cis is not None else 0]
        for block in model.layers:
            c = block.attention.kv_cache
            sig += [block.attention.wqkv.weight.data_ptr(), c.k_cache.data_ptr(), c.pos.data_ptr(), id(c)]
        return tuple(sig)

    def info(self):
        vals = [C.c_int32() for _ in range(4)]
        _lib.check(self.lib.cc_token_info(self.handle, *[C.byref(v) for v in vals]), "cc_token_info")
        return dict(zip(("n_ctas", "n_stages_program", "ring_stages", "smem_bytes"), (v.value for v in vals)))

    def buffer(self, layer: int, which: int) -> torch.Tensor:
        """Payload (bf16 pairs as int32) of a stage output after a decode call -- tests only."""
        off, words = C.c_uint64(), C.c_int32()
        _lib.check(self.lib.cc_token_buffer(self.handle, layer, which, C.byref(off), C.byref(words)), "cc_token_buffer")
        raw = self.scratch[off.value: off.value + 8 * words.value].view(torch.int32).view(-1, 2)
        return raw[:, 0].contiguous().view(torch.bfloat16), raw[:, 1].contiguous()

    def decode(self, token: torch.Tensor, input_pos: torch.Tensor, want_probs: bool = True):
        """token: int32 [1] (or [1, 1]); input_pos: int32 [1].  Returns (token_out int32 [1], probs bf16 [vocab], logits)."""
        assert token.dtype == torch.int32 and input_pos.dtype == torch.int32 and token.is_cuda and input_pos.is_cuda
        _lib.check(self.lib.cc_decode_token(self.handle, token.data_ptr(), input_pos.data_ptr(), self.token_out.data_ptr(),
                                            self.logits.data_ptr(), self.probs.data_ptr() if want_probs else None,
                                            torch.cuda.current_stream().cuda_stream), "cc_decode_token")
        return self.token_out, self.probs, self.logits

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.cc_token_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
