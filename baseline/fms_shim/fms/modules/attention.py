import torch
import torch.nn as nn
import torch.nn.functional as F


class _FusedQKV(nn.Module):
    def __init__(self, emb_dim, nheads, kvheads, hd):
        super().__init__()
        self.splits = [nheads * hd, kvheads * hd, kvheads * hd]
        self.qkv_fused = nn.Linear(emb_dim, sum(self.splits), bias=False)


class MultiHeadAttention(nn.Module):
    def __init__(self, emb_dim, emb_kq, emb_v, nheads, kvheads, p_dropout=None, use_bias=False, position_encoder=None,
                 fused=True):
        super().__init__()
        self.nheads, self.kvheads, self.hd = nheads, kvheads, emb_kq
        self.in_proj = _FusedQKV(emb_dim, nheads, kvheads, emb_kq)
        self.dense = nn.Linear(nheads * emb_v, emb_dim, bias=False)
        self.position_encoder = position_encoder

    def reset_parameters(self):
        for m in (self.in_proj.qkv_fused, self.dense):
            nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02)

    def forward(self, x):
        B, S, _ = x.shape
        q, k, v = self.in_proj.qkv_fused(x).split(self.in_proj.splits, dim=-1)
        q = q.view(B, S, self.nheads, self.hd)
        k = k.view(B, S, self.kvheads, self.hd)
        v = v.view(B, S, self.kvheads, self.hd)
        q, k = self.position_encoder.adjusted_qk(q, k)
        q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        if self.kvheads != self.nheads:
            rep = self.nheads // self.kvheads
            k = k.unsqueeze(2).expand(-1, -1, rep, -1, -1).flatten(1, 2)
            v = v.unsqueeze(2).expand(-1, -1, rep, -1, -1).flatten(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
        return self.dense(o.transpose(1, 2).reshape(B, S, self.nheads * self.hd))
