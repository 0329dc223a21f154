#!/usr/bin/env python
"""bench.py -- decode tokens/sec of the compressed-KV decode path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads = the five BASELINE.json configs (random-init bf16 weights, synthetic token ids):
  1  Llama-3-8B-shape, recent_global, 64 slots, 128-token prompt                       (the reference's CPU plumbing case)
  2  Llama-3-8B-shape, heavy_hitter + SnapKV compression, 1024 slots, 7680-token prompt = "8k ctx"   (default at N = 1)
  3  Llama-3-8B-shape, l2, 512 slots, 4096-token prompt
  4  Qwen2-7B-shape, hybrid (FastGen per-head strategies), 17408 slots allocated / ~2048 live, 16384-token prompt
  5  Llama-3-8B-shape, heavy_hitter with pyramid cache lengths (sum 33128 slots), 7680-token prompt  (default at N > 1,
     tp.py head sharding)
One *step* = one decoded token through every layer.  `value` = decode tokens/s with every input resident in HBM (the
step is ONE persistent kernel, or a CUDA graph of per-layer kernels for config 4, whose only inputs are the previous
token and the position, both on the device); `e2e` = the same through the public decode loop with the token id coming
from pinned host memory and the sampled token read back to the host every step.

`--impl reference` runs the UNMODIFIED reference (installed into baseline/_ref by tools/install_reference.sh, see
DESIGN.md) on the host cores: its own Transformer / KVCache classes and its own decode_one_token, caches in a
synthetic full post-prefill state (the reference's O(S^2) CPU prefill of an 8k prompt takes ~20 minutes).  Without
baseline/_ref it falls back to the port (torch CPU GEMVs + oracle/ for cache and attention) and says so.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LLAMA, QWEN = "Meta-Llama-3-8B-Instruct", "Qwen2-7B-Instruct"
HYBRID_STRATEGIES = [   # ~2048 live slots per head of 17408 allocated
    {"strategy": "window", "recent_window": 0.1176},
    {"strategy": "window_heavy_hitter", "recent_window": 0.0588, "heavy_hitter_frac": 0.0588},
    {"strategy": "full"},
]
CONFIGS = {
    1: dict(model=LLAMA, strategy="recent_global", compressor="recent_global", slots=64.0, pattern="tile", prompt=128, max_new=16,
            what="recent_global, 64 slots/layer, 128-token synthetic prompt"),
    2: dict(model=LLAMA, strategy="heavy_hitter", compressor="heavy_hitter", slots=1024.0, pattern="tile", prompt=7680, max_new=512,
            what="heavy_hitter + SnapKV compression, 1024 slots/layer, 7680-token synthetic prompt, decode @ 8k ctx"),
    3: dict(model=LLAMA, strategy="l2", compressor="l2", slots=512.0, pattern="tile", prompt=4096, max_new=256,
            what="l2, 512 slots/layer, 4096-token synthetic prompt"),
    4: dict(model=QWEN, strategy="hybrid", compressor="full", slots=1.0, pattern="tile", prompt=16384, max_new=1024,
            what="hybrid (FastGen per-head strategies, ~2048 live of 17408 slots per head), 16384-token synthetic prompt"),
    5: dict(model=LLAMA, strategy="heavy_hitter", compressor="heavy_hitter", slots=1024.0, pattern="pyramid", prompt=7680, max_new=512,
            what="heavy_hitter + SnapKV compression, pyramid cache lengths (2036 ... 256, sum 33128 slots), 7680-token "
                 "synthetic prompt, decode @ 8k ctx"),
}
METRIC = "decode tokens/sec @ 8k ctx, 1024-slot heavy-hitter"   # BASELINE.json's metric (configs 2 and 5)
PROMPT_LEN, MAX_NEW = CONFIGS[2]["prompt"], CONFIGS[2]["max_new"]   # (profiles/ scripts)


def workload_config(cfg_id, world):
    """The `config` object both arms print: it names the job (BASELINE config, shapes, how many ways it is split),
    not how an arm executes it -- arm-specific facts go into `notes`."""
    cfg = CONFIGS[cfg_id]
    return {"workload": f"BASELINE config {cfg_id}: {cfg['model']}-shape random-init bf16, {cfg['what']}",
            "global_batch": 1, "seq_len": cfg["prompt"] + cfg["max_new"],
            "parallelism": "none" if world == 1 else f"tp{world}",
            "l2": "no flush needed: the inputs of a step (a rank's weights + live K/V, 2.8-15 GB) exceed the 126 MB L2"}


def metric_name(cfg_id):
    return METRIC if cfg_id in (2, 5) else f"decode tokens/sec (BASELINE config {cfg_id})"


def cache_kwargs(cfg=None):
    cfg = cfg or CONFIGS[2]
    kw = dict(max_cache_length=[cfg["slots"]], cache_bits=None, cache_length_pattern=cfg["pattern"],
              cache_strategy=[cfg["strategy"]], cache_strategy_pattern="tile", feed_long_prompts=False,
              prompt_compression_strategy=[cfg["compressor"]], global_tokens=4, recent_window=10,
              history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)
    if cfg["strategy"] == "hybrid":
        kw["hybrid_strategies"] = [dict(s) for s in HYBRID_STRATEGIES]
    return kw


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------- clocks sampling

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The sampler is started before the warm-up steps (the
    tool needs a few hundred ms to come up; the timed region of a tensor-parallel run is ~0.1-0.5 s) and `mark()`
    brackets the timed region: the summary uses the samples taken inside it, or, if the region was shorter than the
    tool's period, every sample taken under load since the warm-up began."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx, self.marks = [], None, gpu_index, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def mark(self):
        self.marks.append(time.perf_counter())

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        rows, window = self.rows, "warm-up + timed steps"
        if len(self.marks) >= 2:
            inside = [r for r in self.rows if self.marks[0] <= r[0] <= self.marks[-1]]
            if inside:
                rows, window = inside, "timed steps"
        rows = [r[1:] for r in rows]
        sm = [int(r[0]) for r in rows if r[0].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": int(rows[0][1]) if rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(rows), "window": window}


# --------------------------------------------------------------------------- our arm (GPU)

def build_gpu_model(device, world, model_name=LLAMA):
    from cold_compress_b200.model import ModelArgs, Transformer
    cfg = ModelArgs.from_name(model_name)
    cfg.block_size = 32768  # only sizes the RoPE table (random-init model); decode positions run past the prompt
    torch.manual_seed(0)
    with torch.device(device):
        model = Transformer(cfg).to(torch.bfloat16)
    with torch.no_grad():
        gen = torch.Generator(device=device).manual_seed(0)
        for name, p in model.named_parameters():
            if "norm" in name:
                p.fill_(1.0)
            else:
                p.normal_(0.0, 0.02, generator=gen)
    fake = int(os.environ.get("CC_TP_FAKE", "0"))
    if world == 1 and fake > 1:
        # diagnostics (profiles/): rank 0's shard of a `fake`-way split on ONE GPU with the exchanges skipped, i.e. the
        # compute-only time of a tensor-parallel rank; the number it prints is NOT a bench value
        from cold_compress_b200.tp import apply_tp
        os.environ["LOCAL_RANK"], os.environ["LOCAL_WORLD_SIZE"] = "0", str(fake)
        apply_tp(model, group=object())
        for blk in model.layers:
            blk.attention.tp_group = blk.feed_forward.tp_group = None
        os.environ["LOCAL_WORLD_SIZE"] = "1"
    if world > 1:
        from cold_compress_b200.tp import apply_tp, enable_fused_exchange
        apply_tp(model)
        # default: one kernel per token with the all-reduces inside it (stores into every rank's scratch over NVLink).
        # CC_TOKEN_KERNEL=0: per-layer kernels with the exchange fused into the projections; + CC_TP_NCCL=1: NCCL.
        if os.environ.get("CC_TOKEN_KERNEL", "1") == "0" and os.environ.get("CC_TP_NCCL") != "1":
            enable_fused_exchange(model)
    return model.eval()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def live_kv_bytes(model):
    """Algorithmic bytes of the cache stage per token on this rank: live K and V rows, bf16."""
    total = 0
    for b in model.layers:
        c = b.attention.kv_cache
        live = int(c.cache_cts.sum()) if c.cache_cts.numel() > 1 else min(int(c.cache_cts), c.max_cache_length) * c.n_heads
        total += 2 * live * c.head_dim * 2
    return total


def k1_roofline(model, steps, device, pos0):
    """Average duration of the per-layer decode kernel (cc_decode_step) alone: the per-layer launches of one token as
    one CUDA graph (K/V of all layers > L2, plus a 256 MB L2 flush between replays, outside the timed events)."""
    caches = [b.attention.kv_cache for b in model.layers]
    c0 = caches[0]
    Hkv, D = c0.n_heads, c0.head_dim
    G = model.layers[0].attention.n_head // model.layers[0].attention.n_local_heads
    gen = torch.Generator(device=device).manual_seed(7)
    q = (torch.randn(1, Hkv * G, 1, D, device=device, generator=gen) * 1.5).to(torch.bfloat16)
    k = torch.randn(1, Hkv, 1, D, device=device, generator=gen).to(torch.bfloat16)
    v = torch.randn(1, Hkv, 1, D, device=device, generator=gen).to(torch.bfloat16)
    pos = torch.tensor([pos0], dtype=torch.int32, device=device)
    ids = torch.zeros(1, 1, dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for c in caches:  # eager once: plans + output buffers
            c.decode_step(q, k, v, pos, input_ids=ids)
        stream.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            for c in caches:
                c.decode_step(q, k, v, pos, input_ids=ids, overlap_predecessor=True)
        times = []
        for i in range(steps + 3):
            flush.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            graph.replay()
            e1.record(stream)
            e1.synchronize()
            pos += 1
            if i >= 3:
                times.append(e0.elapsed_time(e1) * 1e-3 / len(caches))
    per_launch = statistics.mean(times)
    alg_bytes = live_kv_bytes(model) / len(caches)
    peak, peak_src = peaks()
    achieved = alg_bytes / per_launch / 1e9
    traffic = None
    prof = ROOT / "profiles" / "k1_traffic.json"
    if prof.exists() and c0.max_cache_length == 1024 and Hkv == 8:
        traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "kernel": "cc_decode_kernel (one launch per layer)",
            "us_per_launch": round(per_launch * 1e6, 3), "algorithmic_bytes_per_launch": int(alg_bytes),
            "peak_source": peak_src, "launches_timed": len(times) * len(caches),
            "method": "per-layer launches of one token per CUDA-graph replay, CUDA events per replay, 256 MB L2 flush between replays"}


def run_ours(args):
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    cfg_id = args.config or (2 if world == 1 else 5)
    cfg = CONFIGS[cfg_id]
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    real_stdout = None
    if world > 1:
        # NCCL printf()s its version banner on stdout when NCCL_DEBUG=VERSION is set in the environment; the
        # contract is ONE JSON line on stdout, so point fd 1 at stderr until the result is printed
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("LOCAL_WORLD_SIZE", str(world))
        from cold_compress_b200.tp import maybe_init_dist
        maybe_init_dist()
    dist = torch.distributed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model = build_gpu_model(device, world, cfg["model"])
    ck = cache_kwargs(cfg)
    P, MAXNEW = cfg["prompt"], cfg["max_new"]
    gu.setup_caches(model, SyntheticTokenizer(vocab_size=model.config.vocab_size), device, P + MAXNEW, ck)
    torch.manual_seed(1234)
    prompt = torch.randint(0, model.config.vocab_size, (P,), dtype=torch.int32, device=device)
    with torch.no_grad():
        t0 = time.perf_counter()
        tok, _ = gu.prefill(model, prompt.view(1, -1), torch.arange(P, device=device))
        torch.cuda.synchronize()
        prefill_s = time.perf_counter() - t0
        decoder = gu.GraphedDecoder()
        input_pos = torch.tensor([P], dtype=torch.int32, device=device)
        tok = tok.view(1, 1).clone()
        assert P + args.warmup + 2 * args.steps + 128 < model.config.block_size, "too many steps for the RoPE table"

        def step_device(tok):
            nxt, _ = decoder(model, tok, input_pos)
            input_pos.add_(1)
            return nxt.view(1, 1)

        clocks = ClockSampler(local)
        clocks.__enter__()
        for _ in range(max(3, args.warmup)):
            tok = step_device(tok)
        torch.cuda.synchronize()
        # ---- value: K steps, device-resident inputs, CUDA events, max over ranks
        try:
            barrier()
            clocks.mark()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if os.environ.get("CC_BENCH_NCU"):  # ncu --profile-from-start off: capture only timed steps
                torch.cuda.profiler.start()
            e0.record()
            for _ in range(args.steps):
                tok = step_device(tok)
            e1.record()
            if os.environ.get("CC_BENCH_NCU"):
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
            barrier()
            dev_s = e0.elapsed_time(e1) * 1e-3
            # ---- e2e: token id from pinned host memory in, sampled token back to the host, every step
            host_tok = torch.zeros(1, 1, dtype=torch.int32).pin_memory()
            host_pos = torch.zeros(1, dtype=torch.int32).pin_memory()
            host_out = torch.zeros(1, dtype=torch.int32).pin_memory()
            host_tok.copy_(tok.cpu())
            cur = int(input_pos.item())
            dev_tok = torch.empty(1, 1, dtype=torch.int32, device=device)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                host_pos[0] = cur
                dev_tok.copy_(host_tok, non_blocking=True)
                input_pos.copy_(host_pos, non_blocking=True)
                nxt, _ = decoder(model, dev_tok, input_pos)
                host_out.copy_(nxt.view(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
                host_tok[0, 0] = host_out[0]
                cur += 1
            barrier()
            e2e_s = time.perf_counter() - t0
            clocks.mark()
            input_pos.fill_(cur)
        finally:
            clocks.__exit__(None, None, None)
        times = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = times.tolist()
        token_kernel = model._tplan is not None
        # one token streams every non-embedding weight of this rank plus the live K/V once
        wbytes = sum(p.numel() * p.element_size() for n, p in model.named_parameters() if "tok_embeddings" not in n)
        kvbytes = live_kv_bytes(model)
        peak, peak_src = peaks()
        roof = None
        if rank == 0 and not os.environ.get("CC_BENCH_NCU"):
            if token_kernel:
                # the dominant (only) kernel of a step is cc_token_kernel: one launch per token, timed by the CUDA events
                # around the K timed steps above; algorithmic bytes per launch = this rank's weights + live K/V rows
                per_launch = dev_s / args.steps
                achieved = (wbytes + kvbytes) / per_launch / 1e9
                traffic = None
                prof = ROOT / "profiles" / "r2" / "token_traffic.json"
                if prof.exists() and cfg_id == 2 and world == 1:
                    traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
                roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": traffic,
                        "kernel": "cc_token_kernel (one persistent launch per decoded token: all layers, head, sampling)",
                        "us_per_launch": round(per_launch * 1e6, 1), "algorithmic_bytes_per_launch": int(wbytes + kvbytes),
                        "algorithmic_bytes": {"weights": int(wbytes), "live_kv": int(kvbytes)},
                        "peak_source": peak_src, "launches_timed": args.steps,
                        "method": "CUDA events around the K timed steps (one launch each); inputs of a step (weights + K/V) exceed L2"}
                if args.k1:
                    roof["per_layer_kernel"] = k1_roofline(model, min(args.steps, 32), device, int(input_pos.item()))
            else:
                roof = k1_roofline(model, min(args.steps, 32), device, int(input_pos.item()))
    ours_per_step = model.kernels_per_token() if model.fused_decode else model.config.n_layer
    par = "none"
    if world > 1:
        par = (f"tp{world} (tp.py head / column sharding; " +
               ("all-reduces inside the whole-token kernel, NVLink stores)" if token_kernel else
                "fused NVLink exchange between per-layer kernels)" if getattr(model, "tp_link", None) is not None else "NCCL all-reduce)"))
    elif int(os.environ.get("CC_TP_FAKE", "0")) > 1:
        par = f"DIAGNOSTIC: rank-0 shard of tp{os.environ['CC_TP_FAKE']}, exchanges skipped"
    out = {
        "metric": metric_name(cfg_id), "value": round(args.steps / dev_s, 2), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(dev_s / args.steps * 1e3, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(cfg_id, world),
        "notes": {"parallelism": par,
                  "decode_path": "one persistent kernel per token (cc_decode_token)" if token_kernel else
                                 "per-layer kernels in one CUDA graph (" + str(getattr(model, "token_kernel_reason", "CC_TOKEN_KERNEL=0")) + ")",
                  "prefill_seconds": round(prefill_s, 2)},
        "e2e": {"value": round(args.steps / e2e_s, 2), "unit": "tokens/s", "h2d_bytes_per_step": 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": ours_per_step * args.steps,
        "clocks": clocks.summary(),
    }
    if real_stdout is not None:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    if rank == 0:
        out["roofline"] = roof
        out["step_stream"] = {"bytes_per_token_per_gpu": wbytes + kvbytes, "achieved_gbs": round((wbytes + kvbytes) / (dev_s / args.steps) / 1e9, 1),
                              "frac_of_hbm_peak": round((wbytes + kvbytes) / (dev_s / args.steps) / 1e9 / peak, 4)}
        skip_cpu = world != 1 or os.environ.get("CC_BENCH_NCU") or os.environ.get("CC_TP_FAKE") or os.environ.get("CC_BENCH_NOCPU")
        out["cpu_baseline"] = None if skip_cpu else cpu_baseline(cfg_id, sample_seconds=20.0, warmup=1)
        print(json.dumps(out), flush=True)
    if world > 1:
        # leave without tearing NCCL down: communicators captured in CUDA graphs make
        # destroy_process_group() hang; every rank is past its last collective here
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# --------------------------------------------------------------------------- reference arm (host CPU)

def usable_cores() -> int:
    """Cores this process may use: min(affinity mask, cgroup CPU quota).  torchrun's OMP_NUM_THREADS=1 default is
    NOT honoured here: the reference arm is one process that owns the host."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def _tiled(block, shape, gen):
    """A bf16 tensor of `shape` made of shifted copies of one random block: memcpy-speed init, distinct memory."""
    n = math.prod(shape)
    shift = int(torch.randint(0, block.numel(), (1,), generator=gen))
    return block.roll(shift).repeat(-(-n // block.numel()))[:n].view(*shape).contiguous()


def load_reference():
    """The unmodified reference's modules from baseline/_ref, or None."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "generation_utils.py").exists():
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    import importlib
    mods = {n: importlib.import_module(n) for n in ("model", "cache", "generation_utils")}
    for m in mods.values():
        assert Path(m.__file__).resolve().parent == ref.resolve(), f"{m.__name__} resolved outside baseline/_ref"
    return mods


_REF_STATE = {}


def reference_decoder(cfg_id):
    """(step function, description) running the unmodified reference's decode_one_token on the host."""
    if cfg_id in _REF_STATE:
        return _REF_STATE[cfg_id]
    mods = load_reference()
    if mods is None:
        return None
    cfg = CONFIGS[cfg_id]
    rmodel, rgu = mods["model"], mods["generation_utils"]
    gen = torch.Generator().manual_seed(0)
    block = (torch.randn(8 << 20, generator=gen) * 0.02).to(torch.bfloat16)
    with torch.device("meta"):
        model = rmodel.Transformer.from_name(cfg["model"])
    model = model.to(dtype=torch.bfloat16).to_empty(device="cpu")
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.data = torch.ones(p.shape, dtype=torch.bfloat16) if "norm" in name else _tiled(block, tuple(p.shape), gen)
    model.config.block_size = 32768
    model.eval()
    P, MAXNEW = cfg["prompt"], cfg["max_new"]
    ck = cache_kwargs(cfg)
    tokenizer = None
    if cfg["strategy"] == "hybrid":
        class _Tok:   # the two id lists hybrid caches ask the tokenizer for (tokenizer.py:55-63)
            def special_ids(self):
                return [[0], [1]]

            def punctuation_ids(self):
                return list(range(2, 34))
        tokenizer = _Tok()
    rgu.setup_caches(model, tokenizer, torch.device("cpu"), P + MAXNEW, ck)
    # synthetic full post-prefill state (what the compressed prompt leaves behind): ascending kept positions incl.
    # the global tokens and the recent tail, random K/V and scores
    for layer in model.layers:
        c = layer.attention.kv_cache
        H, L, D = c.k_cache.shape[1], c.k_cache.shape[2], c.k_cache.shape[3]
        live = min(L, P) if cfg["strategy"] != "hybrid" else 2048
        c.k_cache[:, :, :live].copy_(_tiled(block, (1, H, live, D), gen))
        c.v_cache[:, :, :live].copy_(_tiled(block, (1, H, live, D), gen))
        keep = torch.sort(torch.randperm(P - 20, generator=gen)[: live - 20] + 4).values
        pos = torch.cat([torch.arange(4), keep[: live - 20], torch.arange(P - 16, P)]).int()
        c.pos[:, :, :live] = pos
        c.mask[..., :live] = True
        c.cache_cts.fill_(live)
        if hasattr(c, "attn_history_num"):
            c.attn_history_num[:, :, :live].copy_(torch.rand(c.attn_history_num[:, :, :live].shape, generator=gen).to(c.attn_history_num.dtype))
            c.attn_history_denom[:, :, :live] = 1
            c.attn_counter.fill_(1)
        if hasattr(c, "key_norm"):
            c.key_norm.copy_(torch.linalg.vector_norm(c.k_cache.float(), dim=-1).to(c.key_norm.dtype))
        if cfg["strategy"] == "hybrid":
            c.cache_strategies.fill_(1)   # window_heavy_hitter on every head
    from torch.nn.attention import SDPBackend, sdpa_kernel
    state = {"tok": torch.tensor([[17]], dtype=torch.int32), "pos": torch.tensor([P], dtype=torch.int32)}

    def step():
        with torch.no_grad(), sdpa_kernel([SDPBackend.MATH]):   # as decode_n_tokens does (generation_utils.py:193-196)
            nxt, _ = rgu.decode_one_token(model, state["tok"], state["pos"])
        state["tok"] = nxt.view(1, 1)
        state["pos"] += 1

    desc = (f"the unmodified reference (baseline/_ref: its Transformer, KVCache classes and decode_one_token, torch CPU ops), "
            f"{cfg['model']}-shape, {len(model.layers)} layers, caches in a synthetic full post-prefill state")
    _REF_STATE[cfg_id] = (step, desc)
    return _REF_STATE[cfg_id]


def cpu_baseline(cfg_id, sample_seconds=20.0, warmup=1, steps=None):
    """tokens/s of the reference's CPU path on the host cores (all usable cores), on a bounded sample."""
    cores = usable_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    ref = None
    try:
        ref = reference_decoder(cfg_id)
    except Exception as e:   # the installed reference could not be driven: say so and fall back to the port
        sys.stderr.write(f"[bench] reference arm unavailable ({type(e).__name__}: {e}); falling back to the port\n")
    if ref is not None:
        step, desc = ref
        kind = "reference"
    else:
        if cfg_id not in (2, 5):
            return None
        step, desc, kind = port_decoder(), ("port: GEMVs via torch CPU ops as the reference calls them, cache + attention via "
                                            "oracle/cc_oracle.c (OpenMP)"), "port"
    for _ in range(warmup):
        step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and el >= sample_seconds) or el > 240:
            break
    return {"value": round(n / el, 3), "unit": "tokens/s", "cores": cores, "kind": kind,
            "sample": f"{n} full-model decode tokens in {el:.1f} s on {cores} threads; {desc}"}


# ---- the port (only used when baseline/_ref is missing): torch CPU GEMVs + oracle/ cache step

class CpuLayer:
    def __init__(self, cfg, block, seed, slots, prompt):
        from oracle.oracle import OracleCache
        D = cfg["head_dim"]
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)
        kv = cfg["n_kv"] * D
        dim, inter = cfg["dim"], cfg["inter"]
        self.wqkv, self.wo = _tiled(block, (dim + 2 * kv, dim), g), _tiled(block, (dim, dim), g)
        self.w1, self.w3, self.w2 = _tiled(block, (inter, dim), g), _tiled(block, (inter, dim), g), _tiled(block, (dim, inter), g)
        self.n1 = torch.ones(dim, dtype=torch.bfloat16)
        self.n2 = torch.ones(dim, dtype=torch.bfloat16)
        H, L = cfg["n_kv"], slots
        oc = OracleCache("heavy_hitter", H, cfg["n_head"] // H, D, L, 4, 10, 1)
        oc.k_cache.copy_(_tiled(block, (H, L, D), g))
        oc.v_cache.copy_(_tiled(block, (H, L, D), g))
        keep = torch.sort(torch.randperm(prompt - 20, generator=g)[: L - 20] + 4).values
        pos = torch.cat([torch.arange(4), keep[: L - 20], torch.arange(prompt - 16, prompt)]).int()
        oc.pos.copy_(pos.expand(H, L))
        oc.mask.fill_(True)
        oc.cache_cts.fill_(L)
        oc.attn_history_num.copy_(torch.rand(H, L, generator=g, dtype=torch.float64))
        oc.attn_history_denom.fill_(1)
        self.cache = oc

    @staticmethod
    def rmsnorm(x, wt, eps=1e-5):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * wt

    @staticmethod
    def rope(x, fc):
        xs = x.float().reshape(*x.shape[:-1], -1, 2)
        fc = fc.view(1, xs.size(1), 1, xs.size(3), 2)
        re = xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1]
        im = xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]
        return torch.stack([re, im], -1).flatten(3).type_as(x)

    def forward(self, x, freqs, pos):
        cfg = self.cfg
        F = torch.nn.functional
        D, Hq, Hkv = cfg["head_dim"], cfg["n_head"], cfg["n_kv"]
        h = self.rmsnorm(x, self.n1)
        q, k, v = F.linear(h, self.wqkv).split([Hq * D, Hkv * D, Hkv * D], dim=-1)
        q = self.rope(q.view(1, 1, Hq, D), freqs)
        k = self.rope(k.view(1, 1, Hkv, D), freqs)
        y, _, _ = self.cache.decode_step(q.reshape(Hq, D), k.reshape(Hkv, D), v.reshape(Hkv, D).contiguous(), pos)
        x = x + F.linear(y.view(1, 1, Hq * D), self.wo)
        h = self.rmsnorm(x, self.n2)
        return x + F.linear(F.silu(F.linear(h, self.w1)) * F.linear(h, self.w3), self.w2)


_PORT = None


def port_decoder():
    global _PORT
    if _PORT is None:
        cfg = dict(dim=4096, inter=14336, n_head=32, n_kv=8, head_dim=128, n_layer=32, vocab=128256)
        torch.manual_seed(0)
        block = (torch.randn(8 << 20) * 0.02).to(torch.bfloat16)
        layers = [CpuLayer(cfg, block, 100 + i, 1024, 7680) for i in range(cfg["n_layer"])]
        emb = (torch.randn(4096, cfg["dim"]) * 0.02).to(torch.bfloat16)
        out_w = _tiled(block, (cfg["vocab"], cfg["dim"]), torch.Generator().manual_seed(5))
        half = cfg["head_dim"] // 2
        inv = 1.0 / (500000 ** (torch.arange(0, cfg["head_dim"], 2)[:half].float() / cfg["head_dim"]))
        state = {"tok": 17, "pos": 7680}

        def step():
            pos = state["pos"]
            ang = pos * inv
            fr = torch.stack([torch.cos(ang), torch.sin(ang)], -1).to(torch.bfloat16).view(1, half, 2)
            x = emb[state["tok"] % emb.shape[0]].view(1, 1, -1)
            for layer in layers:
                x = layer.forward(x, fr, pos)
            logits = torch.nn.functional.linear(CpuLayer.rmsnorm(x, layers[0].n1), out_w)
            state["tok"] = int(torch.argmax(torch.softmax(logits[0, -1].float(), dim=-1)))
            state["pos"] += 1
        _PORT = step
    return _PORT


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    cfg_id = args.config or (2 if args.gpus == 1 else 5)
    cfg = CONFIGS[cfg_id]
    warm = max(3, args.warmup)   # the same warm-up count as the GPU arm
    base = cpu_baseline(cfg_id, warmup=warm, steps=args.steps)
    if base is None:
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is missing and the port only covers configs 2 and 5"}), flush=True)
        return
    n = int(base["sample"].split()[0])
    out = {"impl": "reference", "metric": metric_name(cfg_id), "value": base["value"], "unit": "tokens/s", "n_gpus": args.gpus,
           "steps": n, "warmup": warm, "ms_per_step": round(1e3 / base["value"], 3),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": workload_config(cfg_id, args.gpus),
           "notes": {"parallelism": "the reference's CPU path on the host cores of rank 0, unsharded"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                    help="BASELINE.json config (default: 2 on one GPU, 5 = pyramid lengths under head sharding on several)")
    ap.add_argument("--k1", action="store_true", help="also time the per-layer decode kernel (cc_decode_step) alone")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
