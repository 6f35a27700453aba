"""MLP speculator (stand-in for ``fms_extras.models.speculator.MLPSpeculator``).

``n_predict`` heads; head i consumes the running state and the embedding of the token i steps ahead:
    state = proj_i(state) * state_weight + emb_i(tok[:, i:i+N]) * emb_weight
    state = GELU(LN_i(state));  logits_i = head_i(state)
with ``state_weight = 0.5 ** (0.5 / n)``, ``emb_weight = sqrt((1 - state_weight^2) * inner_dim / 2)``;
``scale_input`` applies a parameter-free norm / sqrt(2) to the incoming base-model embedding;
``tie_weights`` shares emb / head / ln across heads and proj[1:] (proj[0] maps emb_dim -> inner_dim).
Output: ``[n_predict, B, N, V]`` (indexing used by reference ``train_speculator_utils.py:163-170``).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from fms_fsdp_b200 import ops


class _ScaleShiftRMSNorm(nn.Module):
    """fms LayerNormParameterized(use_mean=False) with optional elementwise scale and shift."""

    def __init__(self, dim, eps=1e-6, affine=True):
        super().__init__()
        self.dim, self.eps, self.affine = dim, eps, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(dim))
            self.bias = nn.Parameter(torch.zeros(dim))

    def reset_parameters(self):
        if self.affine:
            nn.init.ones_(self.weight)
            nn.init.zeros_(self.bias)

    def forward(self, x):
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        y = xf.to(x.dtype)
        if self.affine:
            y = y * self.weight + self.bias
        return y


class MLPSpeculator(nn.Module):
    def __init__(self, emb_dim=4096, inner_dim=0, vocab_size=32000, n_predict=3, tie_weights=False,
                 scale_input=False):
        super().__init__()
        self.n_predict = n_predict
        self.emb_dim = emb_dim
        inner_dim = inner_dim if inner_dim != 0 else emb_dim
        self.inner_dim, self.vsize, self.scale_input = inner_dim, vocab_size, scale_input
        self.emb = nn.ModuleList([nn.Embedding(vocab_size, inner_dim) for _ in range(n_predict)])
        self.proj = nn.ModuleList([nn.Linear(emb_dim if i == 0 else inner_dim, inner_dim, bias=False)
                                   for i in range(n_predict)])
        self.head = nn.ModuleList([nn.Linear(inner_dim, vocab_size, bias=False) for _ in range(n_predict)])
        self.ln = nn.ModuleList([_ScaleShiftRMSNorm(inner_dim) for _ in range(n_predict)])
        if scale_input:
            self.ln0 = _ScaleShiftRMSNorm(emb_dim, affine=False)
        self.state_weight = 0.5 ** (0.5 / n_predict)
        self.emb_weight = math.sqrt((1 - self.state_weight ** 2) * (self.inner_dim / 2))
        if tie_weights:
            assert n_predict > 1, "You cannot tie weights between stages when only 1 exists"
            for i in range(1, n_predict):
                self.emb[i] = self.emb[0]
                self.head[i] = self.head[0]
                self.ln[i] = self.ln[0]
                if i > 1:
                    self.proj[i] = self.proj[1]

    def reset_parameters(self):
        std = 1 / math.sqrt(self.inner_dim)
        for m in self.modules():
            if isinstance(m, (nn.Embedding, nn.Linear)):
                nn.init.trunc_normal_(m.weight, 0, std)
            elif isinstance(m, _ScaleShiftRMSNorm):
                m.reset_parameters()

    def forward(self, state: torch.Tensor, inds: torch.Tensor) -> torch.Tensor:
        """state [B, N, emb_dim] (base-model embeddings), inds [B, N + n_predict - 1] (ground-truth tokens)."""
        out = []
        if self.scale_input:
            state = self.ln0(state) / (2 ** 0.5)
        N = state.size(1)
        for i in range(self.n_predict):
            z = ops.embedding(inds[:, i: i + N].contiguous(), self.emb[i].weight)
            state = torch.add(ops.linear(state, self.proj[i].weight), z, alpha=self.emb_weight / self.state_weight)
            state = F.gelu(self.ln[i](state))
            out.append(ops.linear(state, self.head[i].weight))
        return torch.stack(out, dim=0)

    # ---- sharded-runtime protocol: the whole speculator is one (root) unit; it is driven with
    # ``ShardedModel.forward_backward_custom`` because its inputs are (embeddings, tokens), not tokens.
    def engine_units(self):
        return [], [self]
