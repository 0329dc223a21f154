"""Loading of the committed golden traces (tests/golden/*.npz) and shared comparison helpers."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
ALL_FIXTURES = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz")
                      if not p.stem.startswith(("prefill_", "debug_", "model_")) and "_bits" not in p.stem)
GPU_FIXTURES = ALL_FIXTURES   # every decode trace, incl. the BASELINE cache shapes (cfg*), is verified on the B200


def bf(a: np.ndarray) -> torch.Tensor:
    """uint16 bit patterns -> bf16 tensor."""
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint16)).view(torch.bfloat16)


def bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.uint16).numpy()


class Golden:
    def __init__(self, name: str):
        self.name = name
        self.z = np.load(GOLDEN_DIR / f"{name}.npz")
        self.meta = json.loads(bytes(self.z["meta"]).decode())
        m = self.meta
        self.policy = {"full": "external"}.get(m["strategy"], m["strategy"])
        self.Hkv, self.G, self.D, self.L = m["Hkv"], m["G"], m["D"], m["L"]
        self.T = len(self.z["input_pos"])

    def state(self, prefix: str) -> dict:
        out = {}
        for k in self.z.files:
            if k.startswith(prefix + "_"):
                n = k[len(prefix) + 1:]
                out[n[:-6] if n.endswith("__bf16") else n] = bf(self.z[k]) if n.endswith("__bf16") \
                    else torch.from_numpy(self.z[k])
        return out

    def step(self, t: int) -> dict:
        z, m = self.z, self.meta
        d = dict(q=bf(z["q"][t]), k_new=bf(z["k_new"][t]), v_new=bf(z["v_new"][t]),
                 input_pos=int(z["input_pos"][t]), y=bf(z["y"][t]),
                 fill_idx=torch.from_numpy(z["fill_idx"][t].astype(np.int64)),
                 pos_after=z["pos_after"][t])
        d["probs"] = bf(z["probs"][t]) if "probs" in z.files else None
        d["is_punc"] = self.policy == "hybrid" and int(z["input_id"][t]) in m.get("punctuation", [5])
        return d

    def naive(self) -> bool:
        if self.policy == "hybrid":
            return bool(int(self.z["requires_heavy_hitter"][0]))
        return self.policy == "heavy_hitter"


def bf16_ulp(x: torch.Tensor) -> torch.Tensor:
    """Spacing of bf16 numbers at |x| (fp32 tensor in, fp32 out)."""
    e = torch.floor(torch.log2(x.abs().clamp_min(2.0 ** -126)))
    return torch.pow(2.0, e - 7)


def y_error(y: torch.Tensor, y_ref: torch.Tensor):
    """Per query head: (normwise relative error [Hq], worst elementwise excess [Hq]).

    The elementwise figure is |y - y_ref| divided by the allowance
    max(1 bf16 ulp at that magnitude, 1e-3 * rms of the head's output): two fp32 summation
    orders of the same P.V sum legitimately round to adjacent bf16 values, and elements that
    cancel to ~0 carry the absolute error of the whole sum.  <= 1.0 passes."""
    a, b = y.float(), y_ref.float()
    norm_rel = (a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-30)
    rms = b.pow(2).mean(dim=-1, keepdim=True).sqrt()
    allow = torch.maximum(bf16_ulp(torch.maximum(a.abs(), b.abs())), 1e-3 * rms)
    return norm_rel, ((a - b).abs() / allow).amax(dim=-1)


BF16_QUANTUM = 2.0 ** -8  # relative spacing of bf16: the resolution of the output itself


def check_y(y, y_ref, G: int, flipped_kv_heads=(), what: str = ""):
    """The y bar ("within 1e-3 relative in bf16"), stated so that it is decidable on bf16 outputs:

      * every element within max(1 bf16 ulp, 1e-3 * rms(head)) of the other side -- i.e. the two
        results are the same fp32 value rounded to the same or the ADJACENT bf16 number;
      * per-head normwise relative error <= 2^-8 (one bf16 quantum; a single adjacent-rounding on a
        large element of a 128-vector already costs ~1e-3 normwise, so a tighter per-head bound
        would reject correctly rounded results);
      * callers additionally assert that the MEAN normwise error over a trace is <= 1e-3 (returned).

    A kv head whose bf16 probabilities differ from the other side by an ulp somewhere
    (`flipped_kv_heads`, only possible against the torch-generated golden traces) is held to 2e-2:
    one ulp on a large probability moves y by ~4e-3 * |v|, far more than 1e-3 * |y| when |y| << |v|."""
    norm_rel, excess = y_error(y, y_ref)
    for hq in range(norm_rel.numel()):
        loose = (hq // G) in flipped_kv_heads
        lim_n, lim_e = (2e-2, 20.0) if loose else (BF16_QUANTUM, 1.0)
        assert norm_rel[hq] <= lim_n, f"{what}: q head {hq} normwise rel error {norm_rel[hq]:.2e}"
        assert excess[hq] <= lim_e, f"{what}: q head {hq} element off by {excess[hq]:.2f} x allowance"
    keep = [hq for hq in range(norm_rel.numel()) if (hq // G) not in flipped_kv_heads]
    return float(norm_rel[keep].mean()) if keep else 0.0
