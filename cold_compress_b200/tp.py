"""Tensor parallelism over attention heads and FFN columns (reference tp.py).

Same partitioning as the reference: rank r of ws owns query heads [r*n_head/ws, ...), KV heads
[r*Hkv/ws, ...) (wqkv column-sharded per q/k/v block, tp.py:78-83), wo row-sharded, w1/w3
column- and w2 row-sharded; one sum all-reduce of the attention output and one of the FFN output
per layer (tp.py:134-138, 156-160).  Eviction policies are per KV head, so every rank's cache
kernel works on its own heads with no data-path exchange; the only collective is the
all-reduce after `wo` (NCCL over NVLink; gloo on CPU for the host-logic tests).

One process per GPU under torchrun: LOCAL_RANK / LOCAL_WORLD_SIZE as in the reference.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn

from .model import Attention, FeedForward, Transformer


def _get_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def _get_world_size() -> int:
    return int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))


def is_local():
    return _get_rank() == 0


def maybe_init_dist() -> Optional[int]:
    """Initialise the process group if launched with >1 rank; returns the rank or None."""
    rank, world = _get_rank(), _get_world_size()
    if world < 2:
        return None
    if not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(rank)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", rank))
        else:
            dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    return rank


def _shard(t: torch.Tensor, dim: int, rank: int, world: int) -> torch.Tensor:
    assert t.size(dim) % world == 0, f"cannot shard {tuple(t.shape)} along {dim} over {world} ranks"
    return torch.tensor_split(t, world, dim=dim)[rank]


def _shard_linear(linear: nn.Linear, style: str, rank: int, world: int, splits: List[int] = ()):
    """colwise: split out_features (and bias); rowwise: split in_features.  `splits` shards each block
    of a fused projection separately (q | k | v)."""
    dim = {"colwise": 0, "rowwise": 1}[style]

    def cut(t, d):
        if splits:
            return torch.cat([_shard(part, d, rank, world) for part in t.split(list(splits), dim=d)], dim=d)
        return _shard(t, d, rank, world)

    linear.weight = nn.Parameter(cut(linear.weight, dim).contiguous(), requires_grad=False)
    if linear.bias is not None and style == "colwise":
        linear.bias = nn.Parameter(cut(linear.bias, 0).contiguous(), requires_grad=False)
    if style == "colwise":
        linear.out_features //= world
    else:
        linear.in_features //= world


def _apply_tp_ffn(mlp: FeedForward, rank, world, group):
    _shard_linear(mlp.w1, "colwise", rank, world)
    _shard_linear(mlp.w3, "colwise", rank, world)
    _shard_linear(mlp.w2, "rowwise", rank, world)
    mlp.tp_group = group


def _apply_tp_attn(attn: Attention, rank, world, group):
    kv = attn.n_local_heads * attn.head_dim
    _shard_linear(attn.wqkv, "colwise", rank, world, [attn.dim, kv, kv])
    _shard_linear(attn.wo, "rowwise", rank, world)
    attn.n_head //= world
    attn.dim //= world
    attn.n_local_heads //= world
    attn.head_dim = attn.dim // attn.n_head
    attn.tp_group = group


def apply_tp(model: Transformer, group=None) -> None:
    """Shard a full model in place.  Caches built afterwards get n_local_heads / ws heads
    (config is overwritten like reference tp.py:163-168).  NB: like the reference, `output`
    and the embedding are replicated."""
    rank, world = _get_rank(), _get_world_size()
    if group is None:
        group = dist.group.WORLD
    cfg = model.config
    assert cfg.n_local_heads % world == 0, f"{cfg.n_local_heads} KV heads cannot be split over {world} ranks"
    for block in model.layers:
        _apply_tp_ffn(block.feed_forward, rank, world, group)
        _apply_tp_attn(block.attention, rank, world, group)
    cfg.n_head //= world
    cfg.dim //= world
    cfg.n_local_heads //= world


class FusedExchange:
    """Peer-mapped (NVLink) buffers for the all-reduce fused into the projection kernels (CcTpLink).

    `torch.distributed._symmetric_memory` is used only to allocate and exchange the mappings; the
    data path is this repo's kernels (csrc/gemv.cu: tp_send_pair / tp_send_done / stage_input)."""

    def __init__(self, model: Transformer, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        # the in-kernel collection reads the ranks in pairs (csrc/gemv.cu tp_collect): even worlds only
        assert self.world in (2, 4, 8), f"the fused exchange supports 2, 4 or 8 ranks, not {self.world}"
        self.dim = model.output.in_features
        self.n_ops = 2 * len(model.layers)
        dev = model.output.weight.device
        # u64 [n_ops][world][dim / 2] flag-in-data words = 4 bytes per vector element
        self.buf = symm.empty(self.n_ops * self.world * self.dim, dtype=torch.int32, device=dev)
        self.flags = symm.empty(self.n_ops * self.world, dtype=torch.int32, device=dev)
        self.buf.zero_()
        self.flags.zero_()
        hb = symm.rendezvous(self.buf, group.group_name)
        hf = symm.rendezvous(self.flags, group.group_name)
        torch.cuda.synchronize()
        dist.barrier(group)
        self.peer_buf = torch.tensor(list(hb.buffer_ptrs), dtype=torch.int64, device=dev)
        self.peer_flag = torch.tensor(list(hf.buffer_ptrs), dtype=torch.int64, device=dev)
        self.cta_done = torch.zeros(2 * self.n_ops, dtype=torch.int32, device=dev)
        self.epoch = torch.ones(1, dtype=torch.int32, device=dev)
        self._handles = (hb, hf)

    def tensors(self):
        return [self.peer_buf, self.peer_flag, self.cta_done, self.epoch]

    def meta(self, send_op=-1, recv_op=-1):
        return [self.world, self.rank, self.dim, self.n_ops, send_op, recv_op]


def enable_fused_exchange(model: Transformer, group=None) -> None:
    """Switch the fused decode path of a tensor-parallel model from NCCL all-reduces to the in-kernel exchange."""
    model.tp_link = FusedExchange(model, group)
