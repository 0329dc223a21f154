"""`debug_<strategy>` (KVCacheAnalysis, reference cache.py:1291-1420): a full cache driving a shadow compressed cache
and recording the attention mass the strategy loses.  Trace `debug_hh` was recorded from the reference (with the
one-attribute shim documented in make_golden.py: the class cannot be constructed as shipped).

CPU: the composition of oracle steps (full cache: append + attend; shadow: evict/overwrite, then score with the full
cache's probabilities gathered at the shadow's positions) reproduces the reference's shadow-cache evictions and
attention losses.  GPU: the product class does, through the CUDA kernel."""
import numpy as np
import pytest
import torch

from golden_utils import Golden, bits, check_y
from oracle.oracle import ATTN_NAIVE, PHASE_ATTEND, PHASE_EVICT_FILL, PHASE_SCORE, OracleCache


def _sub_state(state, prefix):
    return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}


def _loss(attn_c):
    return (1 - attn_c.sum(dim=-1)).mean()


def test_oracle_composition_reproduces_reference_attention_loss():
    g = Golden("debug_hh")
    m = g.meta
    Lfull = m["S"] + m["T"]
    init = g.state("init")
    full = OracleCache("external", g.Hkv, g.G, g.D, Lfull, 0, 0, 1, attn_mode=ATTN_NAIVE)
    full.load_state(init)
    comp = OracleCache("heavy_hitter", g.Hkv, g.G, g.D, g.L, m["global_tokens"], m["recent_window"], 1)
    comp.load_state(_sub_state(init, "compressed."))
    losses = []
    for t in range(g.T):
        s = g.step(t)
        slot = torch.tensor([int(full.pos[0].argmin())])            # KVCacheFull: first unfilled slot
        y, _, probs = full.decode_step(s["q"], s["k_new"], s["v_new"], s["input_pos"], fill_idx=slot,
                                       phases=PHASE_EVICT_FILL | PHASE_ATTEND)
        diff = np.abs(bits(probs).astype(np.int32) - bits(s["probs"]).astype(np.int32))
        assert diff.max() <= 2
        flipped = set(np.nonzero(diff.reshape(g.Hkv, -1).any(axis=1))[0].tolist())   # 1-ulp probability differences
        check_y(y, s["y"], g.G, flipped, f"step {t}")
        comp.decode_step(s["q"], s["k_new"], s["v_new"], s["input_pos"], phases=PHASE_EVICT_FILL)
        idx = comp.pos.long().clone()
        idx[idx == -1] = Lfull - 1
        attn_c = probs.gather(1, idx)
        comp.decode_step(s["q"], s["k_new"], s["v_new"], s["input_pos"], phases=PHASE_SCORE, probs_in=attn_c)
        assert np.array_equal(comp.pos.numpy().astype(np.int16), g.z["cpos_after"][t]), f"shadow cache pos differs at step {t}"
        losses.append(_loss(attn_c.unsqueeze(0)))
    want = g.state("final")["attention_losses"][: g.T].float()
    got = torch.stack(losses).float()
    assert torch.allclose(got, want, atol=2.0 ** -7, rtol=0), (got - want).abs().max()


@pytest.mark.gpu
def test_debug_cache_kernel_reproduces_reference_attention_loss():
    from cold_compress_b200.cache import get_cache_constructor
    from device_utils import load_state
    g = Golden("debug_hh")
    m = g.meta
    ctor, relevant = get_cache_constructor("debug_heavy_hitter")
    kw = dict(max_cache_length=g.L, max_seq_length=m["S"] + m["T"], cache_bits=None, global_tokens=m["global_tokens"],
              history_window_size=1, recent_window=m["recent_window"], attn_thresholding=False,
              prompt_compression_strategy="heavy_hitter")
    cache = ctor(1, g.Hkv, g.D, torch.bfloat16, **{k: kw[k] for k in relevant}).cuda()
    init = g.state("init")
    load_state(cache, init)
    load_state(cache.compressed, _sub_state(init, "compressed."))
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(g.T):
        s = g.step(t)
        ip.fill_(s["input_pos"])
        y = cache.decode_step(s["q"].view(1, -1, 1, g.D).cuda(), s["k_new"].view(1, -1, 1, g.D).cuda(),
                              s["v_new"].view(1, -1, 1, g.D).cuda(), ip)
        diff = np.abs(bits(cache._out["probs"]).astype(np.int32) - bits(s["probs"]).astype(np.int32))
        assert diff.max() <= 2
        check_y(y[0, :, 0].cpu(), s["y"], g.G, set(np.nonzero(diff.reshape(g.Hkv, -1).any(axis=1))[0].tolist()), f"step {t}")
        assert np.array_equal(cache.compressed.pos[0].cpu().numpy().astype(np.int16), g.z["cpos_after"][t]), \
            f"shadow cache pos differs at step {t}"
    fin = g.state("final")
    assert int(cache.attention_loss_ctr) == g.T
    got, want = cache.attention_losses[: g.T].float().cpu(), fin["attention_losses"][: g.T].float()
    assert torch.allclose(got, want, atol=2.0 ** -7, rtol=0), (got - want).abs().max()
    stats = cache.compute_statistics(torch.tensor(m["S"] + m["T"]))
    assert abs(stats["attention_loss"] - float(want.mean())) < 1e-2
    assert torch.equal(cache.compressed.k_cache.cpu(), fin["compressed.k_cache"].reshape(cache.compressed.k_cache.shape))
    assert torch.equal(cache.compressed.attn_history_denom.cpu(),
                       fin["compressed.attn_history_denom"].reshape(cache.compressed.attn_history_denom.shape))
