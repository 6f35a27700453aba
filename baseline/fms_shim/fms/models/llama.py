from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from fms.modules.attention import MultiHeadAttention
from fms.modules.embedding import WordEmbedding
from fms.modules.feedforward import GatedLinearUnit
from fms.modules.layernorm import LayerNormParameterized


@dataclass
class LLaMAConfig:
    src_vocab_size: int = 32000
    emb_dim: int = 4096
    norm_eps: float = 1e-5
    nheads: int = 32
    kvheads: int = 0
    nlayers: int = 32
    pad_id: int = -1
    hidden_grow_factor: float = 8 / 3
    multiple_of: int = 256
    activation_fn: str = "swish"
    p_dropout: float = 0.0
    max_expected_seq_len: int = 4096
    ntk_scaling: bool = False
    attn_bias: bool = False
    mlp_bias: bool = False
    tie_heads: bool = False
    rope_theta: float = 10000.0
    linear_config: Optional[dict] = None
    fused_weights: bool = True


class RotaryEmbedding(nn.Module):
    """Interleaved-pair RoPE with a cached [S, dim/2, 2, 2] rotation table (fms convention)."""

    def __init__(self, dim, ratio=10000.0, max_seq_len=2048, ntk_scaling=False):
        super().__init__()
        self.dim, self.ratio, self.max_seq_len, self.ntk_scaling = dim, ratio, max_seq_len, ntk_scaling
        self.cached_freqs = {}

    def _alpha(self, seq_len):
        return 1

    def compute_freqs_cis(self, device, max_seq_len=2048):
        key = str(device)
        if key in self.cached_freqs and self.cached_freqs[key].shape[0] >= max_seq_len:
            return
        inv = 1.0 / (self.ratio ** (torch.arange(0, self.dim, 2, device=device).float() / self.dim))
        ang = torch.outer(torch.arange(max(max_seq_len, self.max_seq_len), device=device).float(), inv)
        self.cached_freqs[key] = torch.stack([ang.cos(), -ang.sin(), ang.sin(), ang.cos()], dim=2).view(*ang.shape, 2, 2)

    def adjusted_qk(self, q, k):
        S = q.size(1)
        self.compute_freqs_cis(q.device, S)
        f = self.cached_freqs[str(q.device)][:S][None, :, None]          # [1,S,1,d/2,2,2]
        def rot(x):
            xr = x.float().view(*x.shape[:-1], -1, 1, 2)                  # [B,S,H,d/2,1,2]
            return (f * xr).sum(-1).flatten(3).type_as(x)
        return rot(q), rot(k)


class LLaMABlock(nn.Module):
    def __init__(self, config: LLaMAConfig, rotary_emb: RotaryEmbedding):
        super().__init__()
        c = config
        kv = c.nheads if c.kvheads == 0 else c.kvheads
        hd = c.emb_dim // c.nheads
        self.ln = LayerNormParameterized(c.emb_dim, eps=c.norm_eps)
        self.ff_ln = LayerNormParameterized(c.emb_dim, eps=c.norm_eps)
        self.attn = MultiHeadAttention(c.emb_dim, hd, hd, c.nheads, kv, position_encoder=rotary_emb)
        self.ff_sub_layer = GatedLinearUnit(c.emb_dim, c.hidden_grow_factor, c.multiple_of)

    def forward(self, x):
        x = x + self.attn(self.ln(x))
        return x + self.ff_sub_layer(self.ff_ln(x))


class LLaMA(nn.Module):
    def __init__(self, config: Optional[LLaMAConfig] = None, **kwargs):
        super().__init__()
        self.config = config if config is not None else LLaMAConfig()
        c = self.config
        self.shared = WordEmbedding(c.src_vocab_size, c.emb_dim, tie_weights=c.tie_heads)
        self.rot_emb = RotaryEmbedding(c.emb_dim // c.nheads, c.rope_theta, c.max_expected_seq_len, c.ntk_scaling)
        self.layers = nn.ModuleList([LLaMABlock(c, self.rot_emb) for _ in range(c.nlayers)])
        self.dec_norm = LayerNormParameterized(c.emb_dim, eps=c.norm_eps)

    def reset_parameters(self):
        for m in self.modules():
            if m is not self and hasattr(m, "reset_parameters") and isinstance(
                    m, (MultiHeadAttention, WordEmbedding, GatedLinearUnit, LayerNormParameterized)):
                m.reset_parameters()

    def forward(self, x, **_):
        h = self.shared(x)
        for layer in self.layers:
            h = layer(h)
        return self.shared(self.dec_norm(h), reverse=True)
