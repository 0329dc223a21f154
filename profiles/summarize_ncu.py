#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/r1_launches.csv profiles/r1_bench_launch_shares.txt "<title>"
    python profiles/summarize_ncu.py full gpurun_out/r1_k1_full.ncu-rep profiles/r1_k1_ncu_full.txt [--traffic-json profiles/k1_traffic.json]

`launches`: per-kernel-name totals/shares of one `ncu --metrics gpu__time_duration.sum` launch list.
`full`: the handful of `--set full` metrics the roofline discussion uses, per captured launch.
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

FULL_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def short(name):
    name = name.replace("void ", "")
    for key in ("cc_decode_kernel", "rmsnorm_qkv_rope", "rmsnorm_ffn_gate", "gemv_residual", "cc_prefill_attn", "cc_"):
        if key in name and key != "cc_":
            return key
    return "torch: " + name.split("(")[0][:60]


def launches(src, dst, title):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        k = short(r["Kernel Name"])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ns)
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write(title + "\n")
        f.write("Per-launch device times are serialized and cold-cache (ncu replay): compare SHARES, not absolutes.\n")
        f.write(f"total {tot / 1e3:.1f} us over {sum(n for n, _ in agg.values())} launches\n")
        for k, (n, t) in agg.items():
            f.write(f"  {k:56s} x{n:4d} {t / 1e3:10.1f} us {100 * t / tot:6.1f}%  avg {t / n / 1e3:8.2f} us\n")
    print(open(dst).read())


def full(src, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none: {src.split('/')[-1]} ({len(data)} captured launches)\n")
        for r in data:
            f.write(f"kernel: {r[col['Kernel Name']][:110]}\n")
        for k in FULL_KEYS:
            if k in col:
                f.write(f"  {k:68s} {units[col[k]]:14s} {'  '.join(r[col[k]] for r in data)}\n")
    print(open(dst).read())
    if traffic_json:
        def val(r, k):
            v, u = float(r[col[k]].replace(",", "")), units[col[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        per = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in data]
        json.dump({"dram_bytes_per_launch": round(sum(per) / len(per)), "launches": len(per), "source": src.split("/")[-1],
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches"},
                  open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
        full(sys.argv[2], sys.argv[3], tj)
