"""Seeded weights that the reference (CPU, build container) and this package (GPU box) can both build in seconds:
every parameter is a window-tiling of one seeded normal block of PRIME elements.  PRIME is prime and larger than any
row count, so no two rows of a matrix see the same window of the block (no duplicated rows, no arg-max ties by
construction), and filling 8 B parameters is a memcpy rather than 8 B normal draws."""
import zlib

import torch

PRIME = 4194301


def seeded_weights(model, seed=0, std=0.02):
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "norm" in name:
                p.fill_(1.0)
                continue
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
            block = (torch.randn(PRIME, generator=g) * std).to(torch.bfloat16)
            n = p.numel()
            p.copy_(block.repeat(-(-n // PRIME))[:n].view(p.shape))
    return model
