"""Tensor parallelism for the FROZEN base model of speculator training (the only place the reference uses TP:
``speculator/train_speculator.py:133-160``).  Megatron layout: column-parallel fused QKV and
gate/up (heads / hidden units split across ranks), row-parallel dense and down projection followed by an
all-reduce, vocab-parallel head followed by an all-gather.  Forward-only (no_grad)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from fms_fsdp_b200.models.llama import LLaMA


def shard_llama_for_tp(model: LLaMA, group) -> LLaMA:
    """Slice the weights of a loaded LLaMA in place; afterwards every block computes on 1/tp of the heads and
    of the MLP hidden units.  The model is flagged so ``EmbedLLaMA.forward`` inserts the collectives."""
    tp, r = dist.get_world_size(group), dist.get_rank(group)
    c = model.config
    hd = c.head_dim
    assert c.nheads % tp == 0 and c.kv_heads % tp == 0 and c.hidden_dim % tp == 0 and c.src_vocab_size % tp == 0, \
        "tp size must divide heads, kv heads, hidden dim and vocab"
    H, KV, F = c.nheads // tp, c.kv_heads // tp, c.hidden_dim // tp
    with torch.no_grad():
        for blk in model.layers:
            w = blk.attn.in_proj.qkv_fused.weight.data
            q, k, v = torch.split(w, [c.nheads * hd, c.kv_heads * hd, c.kv_heads * hd], dim=0)
            blk.attn.in_proj.qkv_fused.weight.data = torch.cat(
                [q[r * H * hd:(r + 1) * H * hd], k[r * KV * hd:(r + 1) * KV * hd], v[r * KV * hd:(r + 1) * KV * hd]]).contiguous()
            blk.attn.dense.weight.data = blk.attn.dense.weight.data[:, r * H * hd:(r + 1) * H * hd].contiguous()
            g, u = blk.ff_sub_layer.wg1_fused.weight.data.chunk(2, dim=0)
            blk.ff_sub_layer.wg1_fused.weight.data = torch.cat([g[r * F:(r + 1) * F], u[r * F:(r + 1) * F]]).contiguous()
            blk.ff_sub_layer.w2.weight.data = blk.ff_sub_layer.w2.weight.data[:, r * F:(r + 1) * F].contiguous()
            blk.attn.nheads, blk.attn.kvheads = H, KV
        V = c.src_vocab_size // tp
        model.shared.head.weight.data = model.shared.head.weight.data[r * V:(r + 1) * V].contiguous()
    model._tp_group, model._tp_size = group, tp
    return model


def tp_all_reduce(x, group):
    dist.all_reduce(x, group=group)
    return x


def tp_all_gather_last(x, group):
    tp = dist.get_world_size(group)
    parts = [torch.empty_like(x) for _ in range(tp)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=-1)
