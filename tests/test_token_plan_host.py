"""Host logic of the whole-token kernel's C ABI that needs no device: argument validation and scratch sizing
(`cc_token_scratch_bytes`, include/cc_b200.h).  The pointers are never dereferenced by these calls."""
import ctypes as C

import pytest

from cold_compress_b200 import _lib

FAKE = 0x10000   # any non-null address: validation only checks for null


def model(n_layers=2, dim=4096, nq=32, nkv=8, D=128, inter=14336, vocab=128256, world=1, rank=0, policy=_lib.POLICY_HEAVY_HITTER,
          L=1024, W=1, attn_mode=_lib.ATTN_NAIVE_BF16):
    layers = (_lib.CcTokenLayer * n_layers)()
    for ly in layers:
        for n in ("attention_norm", "wqkv", "wo", "ffn_norm", "w1", "w3", "w2"):
            setattr(ly, n, FAKE)
        st = ly.cache
        st.policy, st.attn_mode, st.n_kv_heads, st.q_per_kv, st.head_dim, st.max_cache_length = policy, attn_mode, nkv, nq // nkv, D, L
        st.global_tokens, st.recent_window, st.history_window, st.pos_heads, st.cts_heads = 4, 10, W, nkv, 1
        for n in ("k_cache", "v_cache", "pos", "mask", "cache_cts", "attn_history_num", "attn_history_denom", "attn_counter", "key_norm"):
            setattr(st, n, FAKE)
    m = _lib.CcTokenModel()
    m.n_layers, m.dim, m.n_q_heads, m.n_kv_heads, m.head_dim, m.intermediate, m.vocab = n_layers, dim, nq, nkv, D, inter, vocab
    m.layers = C.cast(layers, C.POINTER(_lib.CcTokenLayer))
    for n in ("tok_embeddings", "norm", "output", "freqs_cis"):
        setattr(m, n, FAKE)
    m.world, m.rank = world, rank
    m._keep = layers
    return m


def scratch(m):
    return int(_lib.load().cc_token_scratch_bytes(C.byref(m)))


def test_scratch_grows_with_layers_and_world():
    one, two = scratch(model(n_layers=1)), scratch(model(n_layers=2))
    assert 0 < one < two
    per_layer = two - one
    # a layer's exchange buffers: qkv + y + wo partials + h + gate + w2 partials + x words (8 bytes per bf16 pair), plus
    # the attention exchanges; the partial-vector buffers scale with the world size
    assert per_layer >= 8 * ((32 + 16) * 128 + 32 * 128 + 4096 + 4096 + 14336 + 4096) // 2
    w1, w8 = scratch(model(world=1)), scratch(model(world=8, nq=4, nkv=1, inter=1792))
    assert w8 > 0
    assert scratch(model(world=2)) - w1 == 2 * (2 * 8 * 4096 // 2), "two partial-vector buffers per layer grow by one vector per extra rank"


@pytest.mark.parametrize("kw,needle", [
    (dict(dim=4100), "multiples of 256"),
    (dict(D=96), "head_dim must be 64 or 128"),
    (dict(nq=24, nkv=8), "unsupported q_per_kv"),
    (dict(vocab=128257), "vocabulary rows must be even"),
    (dict(world=9), "bad world"),
    (dict(world=2, rank=2), "bad world"),
    (dict(policy=_lib.POLICY_HYBRID), "not supported by the whole-token kernel"),
    (dict(W=4), "not supported by the whole-token kernel"),
    (dict(policy=_lib.POLICY_L2, attn_mode=_lib.ATTN_NAIVE_BF16), "not supported by the whole-token kernel"),
])
def test_invalid_models_are_refused_with_a_message(kw, needle):
    lib = _lib.load()
    assert scratch(model(**kw)) == 0
    assert needle in lib.cc_last_error().decode(), lib.cc_last_error().decode()


def test_plan_create_checks_its_arguments_before_touching_a_device():
    lib = _lib.load()
    m = model()
    handle = C.c_void_p()
    assert lib.cc_token_plan_create(C.byref(m), None) < 0
    assert lib.cc_token_plan_create(C.byref(m), C.byref(handle)) < 0      # scratch_peers is null
    assert "scratch_peers" in lib.cc_last_error().decode()
    assert lib.cc_decode_token(None, None, None, None, None, None, None) < 0
    assert lib.cc_token_plan_destroy(None) == 0
