"""GPU tests of the whole-token persistent kernel (csrc/decode_token.cu, cc_decode_token):

  * its attention stage alone, fed the reference traces' q / k / v, is held to the per-layer kernel's bar
    (eviction indices, pos, pooled probabilities bit-exact against oracle and reference, y within the bf16 bar);
  * every projection stage of a full token is checked against an fp64 restatement computed from the kernel's own
    stage inputs (<= 1 bf16 ulp);
  * tokens, eviction traces and cache state equal the per-layer-kernel path; CUDA-graph replay equals eager.
"""
import numpy as np
import pytest
import torch

from golden_utils import GPU_FIXTURES, Golden

pytestmark = pytest.mark.gpu


def _eligible(name):
    g = Golden(name)
    return (g.policy in ("recent_global", "l2", "heavy_hitter") and g.meta.get("history_window_size", 1) == 1
            and not g.meta.get("attn_thresholding", False) and g.G in (1, 2, 4, 7, 8))


TRACES = [n for n in GPU_FIXTURES if _eligible(n)]


@pytest.mark.parametrize("name", TRACES)
def test_attention_stage_matches_reference_and_oracle(name):
    from test_gpu_parity import _run_trace
    _run_trace(name, via_token_kernel=True)


SHAPES = [
    ("heavy_hitter", 8, 4, 128, 1024, 1024, 24),   # BASELINE config 2 shape
    ("heavy_hitter", 8, 4, 128, 2036, 2000, 50),   # pyramid layer length: ragged 128-row chunks, fill then evict
    ("heavy_hitter", 4, 7, 128, 556, 556, 24),     # Qwen2-7B grouping, ragged length
    ("heavy_hitter", 1, 4, 128, 1024, 1024, 16),   # one local head (tp.py world size 8): 64 CTAs on one head
    ("heavy_hitter", 2, 8, 64, 300, 120, 30),      # D=64, G=8
    ("l2", 8, 4, 128, 512, 512, 24),               # BASELINE config 3 shape
    ("recent_global", 8, 4, 128, 64, 64, 24),      # BASELINE config 1 shape
    ("recent_global", 8, 4, 128, 4096, 4090, 12),  # two chunks per CTA through the ring
]


@pytest.mark.parametrize("strategy,Hkv,G,D,L,n_filled,steps", SHAPES)
def test_attention_stage_matches_oracle_on_seeded_state(strategy, Hkv, G, D, L, n_filled, steps):
    from test_gpu_parity import _seeded_state_run
    _seeded_state_run(strategy, Hkv, G, D, L, n_filled, steps, via_token_kernel=True)


def _model(strategy="heavy_hitter", L=64, n_layer=2, seed=0, bias=False, n_head=8, n_kv=2, dim=1024, inter=1024, vocab=1000):
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.model import ModelArgs, Transformer
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    cfg = ModelArgs(block_size=512, vocab_size=vocab, n_layer=n_layer, n_head=n_head, dim=dim, n_local_heads=n_kv,
                    intermediate_size=inter, rope_base=500000, attention_bias=bias)
    torch.manual_seed(seed)
    model = Transformer(cfg)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "norm" in name:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            else:
                p.normal_(0.0, 0.05)
    model = model.to(device="cuda", dtype=torch.bfloat16).eval()
    ck = dict(max_cache_length=[float(L)], cache_bits=None, cache_length_pattern="tile", cache_strategy=[strategy],
              cache_strategy_pattern="tile", prompt_compression_strategy=[strategy], global_tokens=4, recent_window=10,
              history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)
    gu.setup_caches(model, SyntheticTokenizer(), torch.device("cuda"), 400, ck)
    return model


def _payload(plan, layer, which, step_no):
    data, tags = plan.buffer(layer, which)
    assert bool((tags == step_no).all()), f"layer {layer} buffer {which}: stale words (tags {tags.unique().tolist()}, want {step_no})"
    return data


@pytest.mark.parametrize("strategy,bias,n_head,n_kv,dim,inter", [
    ("heavy_hitter", False, 8, 2, 1024, 1024),     # Llama grouping (G = 4)
    ("l2", True, 14, 2, 1792, 2560),               # Qwen2 grouping (G = 7, qkv bias), ragged segment tails
    ("recent_global", False, 8, 8, 1024, 2816),    # multi-head attention (G = 1)
])
def test_every_projection_stage_within_one_ulp(strategy, bias, n_head, n_kv, dim, inter):
    import gemv_reference as R
    from cold_compress_b200 import generation_utils as gu
    model = _model(strategy, L=64, n_layer=3, seed=2, bias=bias, n_head=n_head, n_kv=n_kv, dim=dim, inter=inter, vocab=1002)
    torch.manual_seed(3)
    S = 100
    prompt = torch.randint(0, 1002, (S,), dtype=torch.int32, device="cuda")
    with torch.no_grad():
        tok, _ = gu.prefill(model, prompt.view(1, -1), torch.arange(S, device="cuda"))
        pos = torch.tensor([S], dtype=torch.int32, device="cuda")
        tok = tok.view(1, 1).int()
        for step in range(1, 4):
            fused = model.decode_token_fused(tok, pos)
            assert fused is not None, getattr(model, "token_kernel_reason", "")
            plan = model.token_plan()
            nxt, probs = fused[0].clone(), fused[1].clone()
            D = model.layers[0].attention.head_dim
            n_q, n_kv_l = model.layers[0].attention.n_head, model.layers[0].attention.n_local_heads
            x = model.tok_embeddings.weight[int(tok)]
            worst = {}
            for l, block in enumerate(model.layers):
                att, ffn = block.attention, block.feed_forward
                if l > 0:
                    x = _payload(plan, l, 6, step)
                qkv = _payload(plan, l, 0, step)
                want = R.qkv_rope(x, block.attention_norm.weight, block.attention_norm.eps, att.wqkv.weight, att.wqkv.bias,
                                  model.freqs_cis[int(pos)], n_q, n_kv_l, D)
                worst["qkv"] = max(worst.get("qkv", 0), R.ulp_excess(qkv, want))
                y = _payload(plan, l, 1, step)
                wo = _payload(plan, l, 2, step)
                worst["wo"] = max(worst.get("wo", 0), R.ulp_excess(wo, R.proj(att.wo.weight, y)))
                h = _payload(plan, l, 3, step)
                assert torch.equal(h, R.bf(x.float() + wo.float())), "h = x + wo(y) must be one exact bf16 add"
                gate = _payload(plan, l, 4, step)
                worst["gate"] = max(worst.get("gate", 0), R.ulp_excess(gate, R.ffn_gate(h, block.ffn_norm.weight, block.ffn_norm.eps, ffn.w1.weight, ffn.w3.weight)))
                w2 = _payload(plan, l, 5, step)
                worst["w2"] = max(worst.get("w2", 0), R.ulp_excess(w2, R.proj(ffn.w2.weight, gate)))
                xn = _payload(plan, l + 1, 6, step)
                assert torch.equal(xn, R.bf(h.float() + w2.float()))
            logits = plan.logits
            want = R.proj(model.output.weight, R.rmsnorm(xn, model.norm.weight, model.norm.eps))
            worst["head"] = R.ulp_excess(logits, want)
            # the gate multiplies two bf16-rounded factors: each may sit on the adjacent bf16 value, the product on the next
            assert all(v <= (2.0 if k == "gate" else 1.0) for k, v in worst.items()), worst
            # greedy sampling: first index of the largest logit; probabilities = softmax of the bf16 logits
            assert int(nxt) == int(torch.argmax(logits.float()))
            p_ref = torch.softmax(logits.float(), -1)
            assert torch.allclose(probs.float(), p_ref, rtol=2 ** -7, atol=1e-6)
            tok = nxt.view(1, 1)
            pos += 1


@pytest.mark.parametrize("strategy", ["heavy_hitter", "l2", "recent_global"])
def test_token_kernel_equals_per_layer_kernels(strategy):
    """Same model through ONE kernel per token and through the 5-launches-per-layer path: the hidden state drifts by
    fp32 summation order only, eviction decisions and cache bookkeeping stay identical while the logits agree."""
    from cold_compress_b200 import generation_utils as gu
    a, b = _model(strategy, seed=4), _model(strategy, seed=4)
    b.token_kernel = False
    torch.manual_seed(5)
    prompt = torch.randint(0, 1000, (150,), dtype=torch.int32, device="cuda")
    with torch.no_grad():
        ta, _ = gu.prefill(a, prompt.view(1, -1), torch.arange(150, device="cuda"))
        tb, _ = gu.prefill(b, prompt.view(1, -1), torch.arange(150, device="cuda"))
        assert torch.equal(ta, tb)
        pos = torch.tensor([150], dtype=torch.int32, device="cuda")
        tok = ta.view(1, 1).int()
        for t in range(16):
            la = a(tok, pos, is_prefill=False).float().clone()
            assert a._tplan is not None, getattr(a, "token_kernel_reason", "token kernel not used")
            lb = b(tok, pos, is_prefill=False).float()
            assert b._tplan is None
            err, scale = (la - lb).abs().max().item(), lb.abs().max().item()
            assert err <= 0.03 * scale + 0.03, f"step {t}: logits differ by {err} (scale {scale})"
            for xa, xb in zip(a.layers, b.layers):
                ca, cb = xa.attention.kv_cache, xb.attention.kv_cache
                if t < 4:
                    assert torch.equal(ca._fill_idx, cb._fill_idx), f"step {t}: eviction differs"
                assert torch.equal(ca.cache_cts, cb.cache_cts)
                assert int((ca.pos == int(pos)).sum()) == int((cb.pos == int(pos)).sum()) > 0
            tok = torch.randint(0, 1000, (1, 1), dtype=torch.int32, device="cuda")
            pos += 1


def test_token_kernel_graph_replay_equals_eager_and_is_deterministic():
    from cold_compress_b200 import generation_utils as gu
    torch.manual_seed(9)
    prompt = torch.randint(0, 1000, (120,), dtype=torch.int32, device="cuda")
    prefill, eager = gu.compile_funcs(False)
    seqs, states = [], []
    for mode in ("eager", "graph", "eager"):
        m = _model("heavy_hitter", seed=3)
        dec = eager if mode == "eager" else gu.compile_funcs(True)[1]
        seq, probs, stats = gu.generate(m, prompt, prefill, dec, max_new_tokens=24)
        assert m._tplan is not None and m.kernels_per_token() == 1
        assert seq.numel() == 120 + 24 and stats["decode_tokens"] == 24
        seqs.append(seq)
        states.append([(l.attention.kv_cache.pos.clone(), l.attention.kv_cache.attn_history_num.clone(),
                        l.attention.kv_cache.k_cache.clone()) for l in m.layers])
        assert all(abs(float(p.float().sum()) - 1.0) < 2e-2 for p in probs[1:])
    assert torch.equal(seqs[0], seqs[1]) and torch.equal(seqs[0], seqs[2])
    for sa, sb in zip(states[0], states[1]):
        assert all(torch.equal(x, y) for x, y in zip(sa, sb))
    for sa, sb in zip(states[0], states[2]):
        assert all(torch.equal(x, y) for x, y in zip(sa, sb))


def test_unsupported_caches_fall_back_to_per_layer_kernels():
    from cold_compress_b200.token_kernel import unsupported_reason
    m = _model("heavy_hitter", seed=1)
    assert unsupported_reason(m) is None
    m.layers[1].attention.kv_cache.history_window_size = 4
    assert "history window" in unsupported_reason(m)
