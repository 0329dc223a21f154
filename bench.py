#!/usr/bin/env python
"""bench.py -- decode tokens/sec of the compressed-KV decode path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): Llama-3-8B-shape, random-init bf16 weights, synthetic token
ids, cache_strategy=heavy_hitter (SnapKV prompt compression), 1024 slots per layer, 7680-token
prompt + 512 decode positions = 8192 = the model's block size ("8k ctx").  One *step* = one decoded
token through all 32 layers.  `value` = decode tokens/s with every input resident in HBM (the step
is a CUDA graph whose only inputs are the previous token and the position, both on the device);
`e2e` = the same through the public `generate`-style loop with the token id coming from pinned host
memory and the sampled token read back to the host every step.

`--impl reference` times the reference's CPU path for the same workload on the host cores: the
model shell through the same torch CPU ops the reference calls (nn.Linear GEMVs, RMSNorm, RoPE)
and the cache/attention step through oracle/ (the C restatement of cache.py + attention_utils.py).
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

MODEL = "Meta-Llama-3-8B-Instruct"
PROMPT_LEN, MAX_NEW, CACHE_SLOTS = 7680, 512, 1024
METRIC = "decode tokens/sec @ 8k ctx, 1024-slot heavy-hitter"


def cache_kwargs():
    return dict(max_cache_length=[float(CACHE_SLOTS)], cache_bits=None, cache_length_pattern="tile",
                cache_strategy=["heavy_hitter"], cache_strategy_pattern="tile", feed_long_prompts=False,
                prompt_compression_strategy=["heavy_hitter"], global_tokens=4, recent_window=10,
                history_window_size=1, attn_thresholding=False, min_recovery_frac=0.9)


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------- clocks sampling

class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.rows)}


# --------------------------------------------------------------------------- our arm (GPU)

def build_gpu_model(device, world):
    from cold_compress_b200.model import ModelArgs, Transformer
    cfg = ModelArgs.from_name(MODEL)
    cfg.block_size = 16384  # only sizes the RoPE table (random-init model); decode positions start at 7680
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
    if os.environ.get("CC_WARM_MB"):  # experiments: MB of w1/w3 rows the cache step pre-loads into L2 (-1 = none)
        mb = int(os.environ["CC_WARM_MB"])
        model.warm_l2_ffn_bytes = -1 if mb < 0 else mb << 20
    return model.eval()


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def k1_roofline(model, steps, device):
    """Average duration of the fused decode kernel alone: the 32 per-layer launches of one token as
    one CUDA graph (inputs = 134 MB of K/V across layers > L2, plus a 256 MB L2 flush between
    replays, outside the timed events)."""
    caches = [b.attention.kv_cache for b in model.layers]
    c0 = caches[0]
    Hkv, D, L = c0.n_heads, c0.head_dim, c0.max_cache_length
    G = model.config.n_head // model.config.n_local_heads
    gen = torch.Generator(device=device).manual_seed(7)
    q = (torch.randn(1, Hkv * G, 1, D, device=device, generator=gen) * 1.5).to(torch.bfloat16)
    k = torch.randn(1, Hkv, 1, D, device=device, generator=gen).to(torch.bfloat16)
    v = torch.randn(1, Hkv, 1, D, device=device, generator=gen).to(torch.bfloat16)
    pos = torch.tensor([PROMPT_LEN + MAX_NEW - 64], dtype=torch.int32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for c in caches:  # eager once: plans + output buffers
            c.decode_step(q, k, v, pos)
        stream.synchronize()
        with torch.cuda.graph(graph, stream=stream):
            for c in caches:
                c.decode_step(q, k, v, pos, overlap_predecessor=True)  # 32 distinct layers
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
    alg_bytes = 2 * Hkv * L * D * 2  # live K and V rows of the layer, bf16
    peak, peak_src = peaks()
    achieved = alg_bytes / per_launch / 1e9
    traffic = None
    prof = ROOT / "profiles" / "k1_traffic.json"
    if prof.exists():
        traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "kernel": "cc_decode_kernel<128,4,naive>",
            "us_per_launch": round(per_launch * 1e6, 3), "algorithmic_bytes_per_launch": alg_bytes,
            "peak_source": peak_src, "launches_timed": len(times) * len(caches),
            "method": "32 per-layer launches per CUDA-graph replay, CUDA events per replay, 256 MB L2 flush between replays"}


def run_ours(args):
    from cold_compress_b200 import generation_utils as gu
    from cold_compress_b200.tokenizer import SyntheticTokenizer
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
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

    model = build_gpu_model(device, world)
    ck = cache_kwargs()
    gu.setup_caches(model, SyntheticTokenizer(), device, PROMPT_LEN + MAX_NEW, ck)
    torch.manual_seed(1234)
    prompt = torch.randint(0, model.config.vocab_size, (PROMPT_LEN,), dtype=torch.int32, device=device)
    with torch.no_grad():
        t0 = time.perf_counter()
        tok, _ = gu.prefill(model, prompt.view(1, -1), torch.arange(PROMPT_LEN, device=device))
        torch.cuda.synchronize()
        prefill_s = time.perf_counter() - t0
        decoder = gu.GraphedDecoder()
        input_pos = torch.tensor([PROMPT_LEN], dtype=torch.int32, device=device)
        tok = tok.view(1, 1).clone()
        assert PROMPT_LEN + args.warmup + 2 * args.steps + 128 < model.config.block_size, "too many steps for the RoPE table"

        def step_device(tok):
            nxt, _ = decoder(model, tok, input_pos)
            input_pos.add_(1)
            return nxt.view(1, 1)

        for _ in range(max(3, args.warmup)):
            tok = step_device(tok)
        # ---- value: K steps, device-resident inputs, CUDA events, max over ranks
        with ClockSampler(local) as clocks:
            barrier()
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
            input_pos.fill_(cur)
        times = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = times.tolist()
        roof = k1_roofline(model, min(args.steps, 32), device) if rank == 0 and not os.environ.get("CC_BENCH_NCU") else None
    ours_per_step = model.kernels_per_token() if model.fused_decode else model.config.n_layer
    out = {
        "metric": METRIC, "value": round(args.steps / dev_s, 2), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(dev_s / args.steps * 1e3, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{MODEL}-shape random-init bf16, heavy_hitter + SnapKV compression, "
                               f"{CACHE_SLOTS} slots/layer, {PROMPT_LEN}-token synthetic prompt, decode @ 8k ctx",
                   "global_batch": 1, "seq_len": PROMPT_LEN + MAX_NEW, "parallelism": (f"tp{world} ({'all-reduces inside the whole-token kernel (NVLink stores)' if model._tplan is not None else 'fused NVLink exchange' if getattr(model, 'tp_link', None) is not None else 'NCCL all-reduce'})"
                                   if world > 1 else ("none" if not int(os.environ.get("CC_TP_FAKE", "0")) > 1 else
                                                      f"DIAGNOSTIC: rank-0 shard of tp{os.environ['CC_TP_FAKE']}, exchanges skipped")),
                   "l2": "per-step inputs (15 GB of weights + 134 MB of K/V) exceed the 126 MB L2",
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
        # context for tokens/s: one token streams every non-embedding weight of this rank plus the live K/V once
        wbytes = sum(p.numel() * p.element_size() for n, p in model.named_parameters() if "tok_embeddings" not in n)
        kvbytes = sum(2 * c.n_heads * c.max_cache_length * c.head_dim * 2 for c in (b.attention.kv_cache for b in model.layers))
        peak, _src = peaks()
        out["step_stream"] = {"bytes_per_token_per_gpu": wbytes + kvbytes, "achieved_gbs": round((wbytes + kvbytes) / (dev_s / args.steps) / 1e9, 1),
                              "frac_of_hbm_peak": round((wbytes + kvbytes) / (dev_s / args.steps) / 1e9 / peak, 4)}
        out["cpu_baseline"] = (cpu_reference(sample_seconds=20.0, warmup=1)
                               if world == 1 and not os.environ.get("CC_BENCH_NCU") and not os.environ.get("CC_TP_FAKE")
                               and not os.environ.get("CC_BENCH_NOCPU") else None)
        print(json.dumps(out), flush=True)
    if world > 1:
        # leave without tearing NCCL down: communicators captured in CUDA graphs make
        # destroy_process_group() hang; every rank is past its last collective here
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# --------------------------------------------------------------------------- reference arm (host CPU)

class CpuLayer:
    """One transformer block on the host: torch CPU ops for the projections (what the reference
    calls), oracle/ for the cache + attention step."""

    def __init__(self, cfg, block, seed):
        from oracle.oracle import OracleCache
        D = cfg.head_dim
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)

        def w(*shape):  # tiled copies of one random block: memcpy-speed init, distinct memory per weight
            n = math.prod(shape)
            reps = -(-n // block.numel())
            shift = int(torch.randint(0, block.numel(), (1,), generator=g))
            return block.roll(shift).repeat(reps)[:n].view(*shape).contiguous()

        kv = cfg.n_local_heads * D
        self.wqkv, self.wo = w(cfg.dim + 2 * kv, cfg.dim), w(cfg.dim, cfg.dim)
        self.w1, self.w3, self.w2 = w(cfg.intermediate_size, cfg.dim), w(cfg.intermediate_size, cfg.dim), \
            w(cfg.dim, cfg.intermediate_size)
        self.n1 = torch.ones(cfg.dim, dtype=torch.bfloat16)
        self.n2 = torch.ones(cfg.dim, dtype=torch.bfloat16)
        G = cfg.n_head // cfg.n_local_heads
        oc = OracleCache("heavy_hitter", cfg.n_local_heads, G, D, CACHE_SLOTS, 4, 10, 1)
        # synthetic post-prefill state: full cache, ascending kept positions incl. globals and the recent tail
        H, L = cfg.n_local_heads, CACHE_SLOTS
        oc.k_cache.copy_(w(H, L, D))
        oc.v_cache.copy_(w(H, L, D))
        keep = torch.sort(torch.randperm(PROMPT_LEN - 20, generator=g)[: L - 20] + 4).values
        pos = torch.cat([torch.arange(4), keep[: L - 20], torch.arange(PROMPT_LEN - 16, PROMPT_LEN)]).int()
        oc.pos.copy_(pos.expand(H, L))
        oc.mask.fill_(True)
        oc.cache_cts.fill_(L)
        oc.attn_history_num.copy_(torch.rand(H, L, generator=g, dtype=torch.float64))
        oc.attn_history_denom.fill_(1)
        self.cache = oc

    @staticmethod
    def rmsnorm(x, wt, eps):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * wt

    def forward(self, x, freqs, pos):
        from cold_compress_b200.model import apply_rotary_emb
        cfg = self.cfg
        F = torch.nn.functional
        D, Hq, Hkv = cfg.head_dim, cfg.n_head, cfg.n_local_heads
        h = self.rmsnorm(x, self.n1, cfg.norm_eps)
        q, k, v = F.linear(h, self.wqkv).split([Hq * D, Hkv * D, Hkv * D], dim=-1)
        q = apply_rotary_emb(q.view(1, 1, Hq, D), freqs)
        k = apply_rotary_emb(k.view(1, 1, Hkv, D), freqs)
        y, _, _ = self.cache.decode_step(q.reshape(Hq, D), k.reshape(Hkv, D), v.reshape(Hkv, D).contiguous(), pos)
        x = x + F.linear(y.view(1, 1, Hq * D), self.wo)
        h = self.rmsnorm(x, self.n2, cfg.norm_eps)
        return x + F.linear(F.silu(F.linear(h, self.w1)) * F.linear(h, self.w3), self.w2)


_CPU_MODEL = None


def cpu_model():
    global _CPU_MODEL
    if _CPU_MODEL is None:
        from cold_compress_b200.model import ModelArgs, precompute_freqs_cis
        cfg = ModelArgs.from_name(MODEL)
        cfg.block_size = 16384
        torch.manual_seed(0)
        block = (torch.randn(8 << 20) * 0.02).to(torch.bfloat16)
        layers = [CpuLayer(cfg, block, 100 + i) for i in range(cfg.n_layer)]
        emb = (torch.randn(4096, cfg.dim) * 0.02).to(torch.bfloat16)  # only rows actually looked up matter
        out_w = block.repeat(-(-cfg.vocab_size * cfg.dim // block.numel()))[: cfg.vocab_size * cfg.dim].view(
            cfg.vocab_size, cfg.dim)
        freqs = precompute_freqs_cis(cfg.block_size, cfg.head_dim, cfg.rope_base, torch.bfloat16)
        _CPU_MODEL = (cfg, layers, emb, out_w, freqs)
    return _CPU_MODEL


def cpu_decode_token(tok, pos):
    cfg, layers, emb, out_w, freqs = cpu_model()
    x = emb[tok % emb.shape[0]].view(1, 1, -1)
    fr = freqs[pos: pos + 1]
    for layer in layers:
        x = layer.forward(x, fr, pos)
    logits = torch.nn.functional.linear(CpuLayer.rmsnorm(x, layers[0].n1, cfg.norm_eps), out_w)
    return int(torch.argmax(torch.softmax(logits[0, -1].float(), dim=-1)))


def usable_cores() -> int:
    """Cores this process may actually use: min(affinity mask, cgroup CPU quota, OMP_NUM_THREADS)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    if os.environ.get("OMP_NUM_THREADS", "").isdigit():
        n = min(n, int(os.environ["OMP_NUM_THREADS"]))
    return max(1, n)


def cpu_reference(sample_seconds=20.0, warmup=1, steps=None):
    """tokens/s of the reference's CPU path (port): whole-model decode on the host cores."""
    from oracle import oracle
    # usable cores = the affinity mask (a container may see 128 logical CPUs but own 32; oversubscribing
    # torch's intra-op pool beyond that collapses its GEMV throughput by 60x)
    cores = usable_cores()
    torch.set_num_threads(cores)
    cpu_model()
    tok, pos = 17, PROMPT_LEN
    for _ in range(warmup):
        tok = cpu_decode_token(tok, pos)
        pos += 1
    n, t0 = 0, time.perf_counter()
    while True:
        tok = cpu_decode_token(tok, pos)
        pos += 1
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and el >= sample_seconds) or el > 240:
            break
    return {"value": round(n / el, 3), "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"{n} full-model decode tokens ({MODEL}-shape, 32 layers, 1024-slot heavy_hitter caches in a "
                      f"synthetic full post-prefill state) in {el:.1f} s; GEMVs via torch CPU ops as the reference "
                      f"calls them, cache+attention via oracle/cc_oracle.c with {oracle.threads()} OpenMP threads"}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    base = cpu_reference(warmup=max(1, min(args.warmup, 3)), steps=args.steps)
    n = int(base["sample"].split()[0])
    out = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "tokens/s", "n_gpus": args.gpus,
           "steps": n, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": round(1e3 / base["value"], 3),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": f"{MODEL}-shape random-init bf16, heavy_hitter, {CACHE_SLOTS} slots/layer, decode @ 8k ctx "
                                  "(reference CPU path, host cores)", "global_batch": 1,
                      "seq_len": PROMPT_LEN + MAX_NEW, "parallelism": "none"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
