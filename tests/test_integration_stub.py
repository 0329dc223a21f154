"""The ctypes stub INTEGRATION.md shows a reference maintainer (section 2) is executed verbatim: it must drive
cc_decode_step on a cache object that has the reference's buffer names, and give the product API's result."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _stub_source():
    text = (ROOT / "INTEGRATION.md").read_text()
    blocks = re.findall(r"```python\n(.*?)```", text, re.S)
    stub = [b for b in blocks if "def fused_decode_step" in b]
    assert len(stub) == 1
    return stub[0]


def test_stub_struct_matches_the_header_layout():
    """CPU: the documented ctypes structs have the sizes of the C structs (every field present, in order)."""
    import ctypes as C
    import os
    from cold_compress_b200 import _lib
    os.environ["CC_B200_LIB"] = str(_lib.SO_PATH)
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    assert C.sizeof(ns["CcDecodeArgs"]) == C.sizeof(_lib.CcDecodeArgs)
    assert C.sizeof(ns["CcCacheState"]) == C.sizeof(_lib.CcCacheState)
    assert [n for n, _ in ns["CcCacheState"]._fields_] == list(_lib._STATE_INTS + _lib._STATE_PTRS)
    assert ns["CcDecodeArgs"].is_punc_dev.offset == _lib.CcDecodeArgs.is_punc_dev.offset


@pytest.mark.gpu
def test_stub_drives_cc_decode_step():
    import os
    from cold_compress_b200 import _lib
    from device_utils import make_device_cache, synthetic_fill
    os.environ["CC_B200_LIB"] = str(_lib.SO_PATH)
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md", "exec"), ns)
    Hkv, G, D, L = 8, 4, 128, 256
    a = make_device_cache("heavy_hitter", Hkv, D, L)
    b = make_device_cache("heavy_hitter", Hkv, D, L)
    for c in (a, b):
        synthetic_fill(c, L, seed=5)
    gen = torch.Generator().manual_seed(1)
    ip = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(6):
        q = (torch.randn(1, Hkv * G, 1, D, generator=gen) * 1.5).to(torch.bfloat16).cuda()
        k = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        v = torch.randn(1, Hkv, 1, D, generator=gen).to(torch.bfloat16).cuda()
        ip.fill_(L + t)
        y_api = a.decode_step(q, k, v, ip).clone()
        y_stub = ns["fused_decode_step"](b, q, k, v, ip)
        torch.cuda.synchronize()
        assert torch.equal(y_api.view(-1), y_stub.view(-1))
        for n in ("pos", "k_cache", "attn_history_num", "attn_history_denom", "cache_cts", "mask"):
            assert torch.equal(getattr(a, n), getattr(b, n)), n
