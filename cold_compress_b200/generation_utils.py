"""Generation runtime (reference generation_utils.py): cache set-up from CLI kwargs, prefill, the
decode loop, statistics.  Same entry points and argument meaning; `compile_funcs(True)` returns a
decode step captured in a hand-built CUDA graph (`GraphedDecoder`) instead of `torch.compile` --
there is no tracing compiler anywhere on the path.
"""
from __future__ import annotations

import argparse
import itertools
import time
from pathlib import Path
from typing import Optional, Tuple

import torch
import yaml

from .model import Transformer, find_multiple

default_device = "cuda" if torch.cuda.is_available() else "cpu"


# --------------------------------------------------------------------------- CLI / reporting helpers

def snake_to_capitalized(s):
    return " ".join(w.capitalize() for w in s.split("_"))


def print_stats(stats_dict):
    flat, layered = {}, {}
    for key, value in stats_dict.items():
        head, _, tail = key.rpartition("_")
        if head and tail.isdigit():
            layered.setdefault(snake_to_capitalized(head), []).append((int(tail), value))
        else:
            flat[snake_to_capitalized(key)] = value
    for key, value in flat.items():
        print(f"{key}: {value:.02f}")
    for stat in sorted(layered):
        row = ", ".join(f"{layer}={value:.02f}" for layer, value in sorted(layered[stat]))
        print(f"{stat} By Layer: {row}")


def add_generation_arguments(parser: argparse.ArgumentParser):
    group = parser.add_argument_group("generation_args")
    group.add_argument("--checkpoint_path", type=Path,
                       default=Path(__file__).resolve().parent / "checkpoints/meta-llama/Meta-Llama-3-8B-Instruct/model.pth",
                       help="Model checkpoint path.")
    group.add_argument("--profile", type=Path, default=None, help="Profile path.")
    group.add_argument("--compile", action="store_true",
                       help="Capture the decode step in a CUDA graph (the reference's torch.compile switch).")
    group.add_argument("--device", type=str, default=default_device, help="Device to use")
    group.add_argument("--attn_top_k", type=float, default=1.0, help="Kept for CLI compatibility; must be 1.0.")


def merge_cache_config(args):
    """Overlay cache_configs/<name>.yaml on the argparse namespace (reference generation_utils.py:79-91)."""
    if not args.cache_config:
        return args
    name = args.cache_config if args.cache_config.endswith(".yaml") else args.cache_config + ".yaml"
    args.cache_config = name
    path = Path(__file__).parent / "cache_configs" / name
    assert path.exists(), f"Cache config file {path} does not exist."
    with open(path) as f:
        overlay = yaml.safe_load(f)
    return argparse.Namespace(**{**vars(args), **overlay})


def compute_max_seq_length(model, prompt_lens, target_lens, max_new_tokens: int):
    longest_prompt = max(len(p) for p in prompt_lens)
    longest_target = 0 if target_lens is None else max(len(t) for t in target_lens)
    max_seq_length = longest_prompt + max(max_new_tokens, longest_target)
    if max_seq_length > model.config.block_size:
        print(f"Warning: max_seq_length {max_seq_length} exceeds the model's block size "
              f"{model.config.block_size}; clamping.")
        max_seq_length = model.config.block_size
    print(f"Maximum context length of {max_seq_length} tokens.")
    return longest_prompt, max_seq_length


def device_sync(device):
    if "cuda" in str(device):
        torch.cuda.synchronize(device)
    elif "cpu" not in str(device) and "mps" not in str(device):
        print(f"device={device} is not yet suppported")


# --------------------------------------------------------------------------- sampling / steps

def greedy(logits, next_token):
    probs = torch.nn.functional.softmax(logits[0, -1], dim=-1)
    if next_token is None:
        return torch.argmax(probs, keepdim=True).to(dtype=torch.int), probs
    return next_token, probs


def prefill(model: Transformer, x: torch.Tensor, input_pos: torch.Tensor, next_token: torch.Tensor = None,
            **sampling_kwargs):
    # no S x S bool mask: causality is applied inside the attention (the reference builds the
    # mask on the host, generation_utils.py:153-158)
    logits = model(x, input_pos, mask=None, is_prefill=True)
    return greedy(logits, next_token)


def decode_one_token(model: Transformer, x: torch.Tensor, input_pos: torch.Tensor, next_token: torch.Tensor = None,
                     attn_top_k: float = 1, **sampling_kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
    fused = model.decode_token_fused(x, input_pos) if attn_top_k in (None, 1, 1.0) and hasattr(model, "decode_token_fused") else None
    if fused is not None:   # ONE kernel: the whole layer stack, the head, arg-max and softmax (csrc/decode_token.cu)
        tok, probs = fused
        return (tok if next_token is None else next_token), probs
    logits = model(x, input_pos, is_prefill=False, attn_top_k=attn_top_k)
    return greedy(logits, next_token=next_token)


class GraphedDecoder:
    """decode_one_token captured once in a CUDA graph and replayed per token.

    Static inputs: the token id and `input_pos` tensors the caller passes on the first call are
    kept as the graph's inputs; later calls copy into them.  The per-layer kernels take
    `input_pos` from device memory, so one graph serves every position."""

    def __init__(self):
        self.graph = None
        self.warm = False

    def _capture(self, model, x, input_pos, attn_top_k):
        self.x = x.clone()
        self.pos = input_pos.clone()
        self.model = model
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            # capture only records: the cache state is untouched until the first replay
            with torch.cuda.graph(self.graph, stream=stream):
                self.tok, self.probs = decode_one_token(model, self.x, self.pos, attn_top_k=attn_top_k)
        torch.cuda.current_stream().wait_stream(stream)

    def __call__(self, model, x, input_pos, next_token=None, attn_top_k=1, **kw):
        if not self.warm:
            # first token eagerly: library handles, launch plans and output buffers get created
            # outside of stream capture (it is a real decode step, nothing is thrown away)
            self.warm = True
            return decode_one_token(model, x, input_pos, next_token=next_token, attn_top_k=attn_top_k)
        if self.graph is None:
            self._capture(model, x, input_pos, attn_top_k)
        self.x.copy_(x.view_as(self.x))
        self.pos.copy_(input_pos)
        self.graph.replay()
        return (self.tok if next_token is None else next_token), self.probs


def decode_n_tokens(model: Transformer, cur_token: torch.Tensor, input_pos: torch.Tensor, decode_one_token: callable,
                    num_new_tokens: int, terminator_ids: Optional[list] = None, attn_top_k: float = 1,
                    prefix: Optional[torch.Tensor] = None, **sampling_kwargs):
    new_tokens, new_probs = [], []
    for i in range(num_new_tokens):
        forced = prefix is not None and i < len(prefix)
        tok, prob = decode_one_token(model, cur_token, input_pos, next_token=prefix[i].view(1) if forced else None,
                                     attn_top_k=attn_top_k, **sampling_kwargs)
        new_tokens.append(tok.clone())
        new_probs.append(prob.clone())
        if terminator_ids and not forced and tok in terminator_ids:  # host sync, only if asked for
            break
        input_pos += 1
        cur_token = tok.view(1, -1)
    return new_tokens, new_probs


def model_forward(model, x, input_pos):
    return model(x, input_pos)


# --------------------------------------------------------------------------- per-layer patterns

def normalize_cache_length(max_cache_length: float, max_seq_length: int, multiple_of: int = 8) -> int:
    """Fraction or absolute -> absolute slots, rounded up to a multiple of 8, capped at the sequence length."""
    if 0 < max_cache_length <= 1:
        slots = round(max_seq_length * max_cache_length)
    else:
        assert int(max_cache_length) == max_cache_length
        slots = int(max_cache_length)
        if slots > max_seq_length:
            print(f"FYI: max_cache_length ({slots}) is greater than max_seq_length ({max_seq_length}). "
                  f"Setting to {max_seq_length}")
            slots = max_seq_length
    return min(find_multiple(slots, multiple_of), max_seq_length)


def apply_pyramid_pattern(max_cache_length: int, max_seq_length: int, model_n_layer: int, decreasing: bool = True,
                          min_cache_length: int = 256):
    """PyramidKV budget schedule (arXiv 2406.02069, beta = 14): a linear ramp over layers from
    total/(n*beta) to 2*total/n slots, floored at min(256, max_cache_length) with the excess taken
    evenly from the layers above the floor; reversed for 'pyramid' (reference generation_utils.py:279-321)."""
    beta = 14
    floor = min(min_cache_length, max_cache_length)
    n = model_n_layer
    total = max_cache_length * n
    lo, hi = total / (n * beta), 2 * total / n
    step = (hi - lo) / n
    ramp = [lo + step * layer for layer in range(n - 1)] + [hi]
    lens = [normalize_cache_length(int(v), max_seq_length) for v in ramp]
    short = [max(0, floor - v) for v in lens]
    n_short = sum(1 for s in short if s > 0)
    lens = [max(v, floor) for v in lens]
    if n_short < n:
        take = sum(short) // (n - n_short)
        lens = [max(floor, v - take) if v > floor else v for v in lens]
    if decreasing:
        lens = lens[::-1]
        assert lens[-1] < lens[0], "Cache lengths should be decreasing."
    else:
        assert lens[0] < lens[-1], "Cache lengths should be increasing."
    return lens


def apply_pattern(pattern: list, out_size: int, extension_strategy: str = "tile", max_seq_length: int = None):
    assert extension_strategy in {"tile", "repeat", "pyramid", "funnel"}
    assert out_size % len(pattern) == 0, f"{len(pattern)} must divide the number of layers ({out_size})."
    reps = out_size // len(pattern)
    if extension_strategy in {"funnel", "pyramid"}:
        assert len(pattern) == 1, "Funnel and pyramid patterns must have a single element."
        return apply_pyramid_pattern(pattern[0], max_seq_length, out_size, decreasing=extension_strategy == "pyramid")
    if extension_strategy == "tile":
        return [item for item in pattern for _ in range(reps)]
    return list(pattern) * reps


def setup_caches(model: Transformer, tokenizer, device, max_seq_length: int, cache_kwargs: dict = None) -> dict:
    """Expand the CLI-level cache kwargs to per-layer values and build the caches
    (reference generation_utils.py:324-388)."""
    ck = cache_kwargs
    n_layer = model.config.n_layer
    ck["max_seq_length"] = max_seq_length
    ck["max_cache_length"] = apply_pattern(
        [normalize_cache_length(v, max_seq_length) for v in ck["max_cache_length"]], n_layer,
        ck["cache_length_pattern"], max_seq_length)
    assert len(ck["cache_strategy"]) == len(ck["prompt_compression_strategy"]), \
        "You must specify a prompt_compression_strategy for each cache_strategy."
    for key in ("cache_strategy", "prompt_compression_strategy"):
        ck[key] = apply_pattern(ck[key], n_layer, ck["cache_strategy_pattern"])
    rw = ck["recent_window"]
    if not isinstance(rw, list):
        ck["recent_window"] = [max(1, int(rw * L)) if rw <= 1 else max(1, min(rw, L)) for L in ck["max_cache_length"]]
    assert ck["global_tokens"] <= min(ck["max_cache_length"]), "Global tokens must be less than max_cache_length."
    if ck["cache_strategy"][0] == "hybrid":
        ck["token_ids"] = {"special": tokenizer.special_ids(), "punctuation": tokenizer.punctuation_ids()}
    with torch.device(device):
        model.setup_caches(max_batch_size=1, **ck)
    return ck


def reset_caches(model: Transformer):
    model.reset_caches()


def get_cache_stats(model: Transformer, prompt_len: int, gen_len: int):
    return model.get_cache_stats(prompt_len, gen_len)


# --------------------------------------------------------------------------- generate

@torch.no_grad()
def generate(model: Transformer, prompt: torch.Tensor, prefill: callable, decode_one_token: callable,
             max_new_tokens: int, next_tokens: Optional[torch.Tensor] = None, terminator_ids: Optional[list] = None,
             feed_long_prompts: bool = False, decode_first_token: bool = False, attn_top_k: float = 1,
             **sampling_kwargs):
    """Prefill the prompt, then decode; returns (sequence, per-token probabilities, perf stats)
    with the reference's accounting (generation_utils.py:399-531)."""
    prompt_length = prompt.size(0)
    device, dtype = prompt.device, prompt.dtype
    min_cache = model.min_cache_length()
    prefix = None
    # long prompts fed token by token, and the edge case |prompt| == cache length
    if (feed_long_prompts and prompt_length > min_cache - 1) or prompt_length == min_cache:
        cut = min_cache - 1
        prompt, prefix = prompt[:cut], prompt[cut:]
        max_new_tokens += len(prefix)
        prompt_length = cut
    if decode_first_token:
        prompt, prefix = prompt[:-1], prompt[-1:]
        max_new_tokens += 1
        prompt_length -= 1
    if next_tokens is not None:  # teacher forcing (perplexity)
        max_new_tokens = len(next_tokens)
        next_token, prefix = next_tokens[0].view(1), next_tokens[1:]
    elif prefix is not None:
        next_token, prefix = prefix[0].view(1), prefix[1:]
    else:
        next_token = None
    seq = torch.full((prompt_length + max_new_tokens,), -1, dtype=dtype, device=device)
    seq[:prompt_length] = prompt
    input_pos = torch.arange(0, prompt_length, device=device)

    device_sync(device)
    t0 = time.perf_counter()
    tok, tok_probs = prefill(model, prompt.view(1, -1), input_pos, next_token=next_token, **sampling_kwargs)
    device_sync(device)
    t1 = time.perf_counter()
    tok, tok_probs = tok.clone(), tok_probs.clone()
    seq[prompt_length] = tok
    input_pos = torch.tensor([prompt_length], device=device, dtype=torch.int)
    gen_tokens, gen_probs = decode_n_tokens(model, tok.view(1, -1), input_pos, decode_one_token, max_new_tokens - 1,
                                            terminator_ids=terminator_ids, prefix=prefix, attn_top_k=attn_top_k,
                                            **sampling_kwargs)
    device_sync(device)
    t2 = time.perf_counter()

    decode_tokens = len(gen_tokens) + 1  # +1: the token produced by prefill
    prefill_s, decode_s, total_s = t1 - t0, t2 - t1, t2 - t0
    perf_stats = {
        "prefill_tokens": prompt_length, "decode_tokens": decode_tokens,
        "prefill_toks_per_sec": prompt_length / prefill_s, "decode_toks_per_sec": decode_tokens / decode_s,
        "total_toks_per_sec": decode_tokens / total_s, "total_seconds": total_s, "prefill_seconds": prefill_s,
        "decode_seconds": decode_s, "decode_seconds_frac_of_total": decode_s / total_s,
        "memory_used_gb": torch.cuda.max_memory_reserved() / 1e9 if torch.cuda.is_available() else 0.0,
    }
    if gen_tokens:
        seq[prompt_length + 1: prompt_length + 1 + len(gen_tokens)] = torch.cat(gen_tokens)
    if -1 in seq:
        seq = seq[: torch.where(seq == -1)[0][0]]
    return seq, [tok_probs] + gen_probs, perf_stats


# --------------------------------------------------------------------------- model loading

def load_model(checkpoint_path, device, precision, use_tp):
    with torch.device("meta"):
        model = Transformer.from_name(checkpoint_path.parent.name)
    if "int8" in str(checkpoint_path) or "int4" in str(checkpoint_path):
        raise NotImplementedError("weight-only quantised checkpoints are outside this build (bf16 weights only)")
    checkpoint = torch.load(str(checkpoint_path), mmap=True, weights_only=True)
    if "model" in checkpoint and "stories" in str(checkpoint_path):
        checkpoint = checkpoint["model"]
    model.load_state_dict(checkpoint, assign=True)
    if use_tp:
        from .tp import apply_tp
        print("Applying tensor parallel to model ...")
        apply_tp(model)
    return model.to(device=device, dtype=precision).eval()


def get_model_size(model):
    """Bytes of every parameter/buffer outside embeddings (reference generation_utils.py:569-575)."""
    total = 0
    for _, child in model.named_children():
        if not isinstance(child, torch.nn.Embedding):
            total += sum(p.numel() * p.dtype.itemsize for p in itertools.chain(child.parameters(), child.buffers()))
    return total


def compile_funcs(compile=True):
    """(prefill, decode_one_token).  compile=True: decode runs as a captured CUDA graph."""
    if compile:
        return prefill, GraphedDecoder()
    return prefill, decode_one_token
