"""Loading of the prefill golden vectors (tests/golden/prefill_*.npz) and shared comparison helpers."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from golden_utils import GOLDEN_DIR, bf, bf16_ulp

PREFILL_FIXTURES = sorted(p.stem for p in GOLDEN_DIR.glob("prefill_*.npz"))


class PrefillGolden:
    def __init__(self, name: str):
        z = np.load(GOLDEN_DIR / f"{name}.npz")
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.t = {}
        for k in z.files:
            if k == "meta":
                continue
            self.t[k[:-6] if k.endswith("__bf16") else k] = bf(z[k]) if k.endswith("__bf16") else torch.from_numpy(z[k])
        m = self.meta
        self.Hkv, self.G, self.D, self.S, self.L = m["Hkv"], m["G"], m["D"], m["S"], m["L"]
        self.naive = bool(m["return_attn"])

    def qkv(self, device="cpu"):
        return tuple(self.t[n].unsqueeze(0).to(device) for n in ("q", "k", "v"))


def ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a - b| in units of the bf16 spacing at max(|a|, |b|) (fp32 out)."""
    a, b = a.float(), b.float()
    return (a - b).abs() / bf16_ulp(torch.maximum(a.abs(), b.abs()))


def assert_close_bf16(got, want, what, max_ulp=1.0, min_equal=0.98, floor=0.0):
    """bf16 results of the same fp32 quantity summed in another order: equal, or adjacent bf16 values on a
    small fraction.  `floor`: values this small (absolute) may differ by `max_ulp` spacings of `floor`."""
    g, w = got.float().cpu(), want.float().cpu()
    d = (g - w).abs() / bf16_ulp(torch.maximum(torch.maximum(g.abs(), w.abs()), torch.tensor(floor)))
    frac_equal = float((g == w).float().mean())
    assert float(d.max()) <= max_ulp, f"{what}: worst difference {float(d.max()):.2f} bf16 ulp (limit {max_ulp})"
    assert frac_equal >= min_equal, f"{what}: only {frac_equal:.4f} of the elements bit-equal (limit {min_equal})"
    return frac_equal


def keep_sets_equal_up_to_ties(keep_got, keep_want, priority_want, what):
    """Same kept tokens, except among tokens whose reference priority equals (or is adjacent to) the cut-off."""
    H = keep_want.shape[0]
    for h in range(H):
        a, b = set(keep_got[h].tolist()), set(keep_want[h].tolist())
        if a == b:
            continue
        pr = priority_want[h].float()
        cutoff = pr[torch.tensor(sorted(b))].min()
        for j in a ^ b:
            d = float((pr[j] - cutoff).abs() / bf16_ulp(torch.maximum(pr[j].abs(), cutoff.abs())))
            assert d <= 2.0, f"{what}: head {h} token {j} kept/dropped differently, priority {float(pr[j])} vs cut-off {float(cutoff)}"


def check_prefill_y(y, y_ref, what, strict_frac=0.995):
    """y rows [N, D] of the kernel against the reference's.  Same bar as golden_utils.check_y (equal or adjacent
    bf16 values, normwise <= 2^-8, mean normwise <= 1e-3) for at least `strict_frac` of the rows; a row whose
    probabilities differ from the reference's by a bf16 ulp somewhere (a logit on a rounding boundary under a
    different dot-product order, not observable from y alone) is held to the documented loose bound 2e-2 / 20x."""
    from golden_utils import BF16_QUANTUM, y_error
    norm_rel, excess = y_error(y.cpu(), y_ref.cpu())
    strict = (norm_rel <= BF16_QUANTUM) & (excess <= 1.0)
    assert float(strict.float().mean()) >= strict_frac, f"{what}: only {float(strict.float().mean()):.4f} of the rows meet the strict bar"
    assert float(norm_rel.max()) <= 2e-2 and float(excess.max()) <= 20.0, \
        f"{what}: worst row normwise {float(norm_rel.max()):.2e}, elementwise {float(excess.max()):.1f}x allowance"
    assert float(norm_rel.mean()) <= 1e-3, f"{what}: mean normwise error {float(norm_rel.mean()):.2e}"
    return float(norm_rel.mean())
