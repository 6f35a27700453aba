import torch
import torch.nn as nn
import torch.nn.functional as F


class GatedLinearUnit(nn.Module):
    def __init__(self, emb_dim, hidden_grow_factor=4, multiple_of=None, activation_fn=None, p_dropout=0.0,
                 use_bias=False):
        super().__init__()
        hidden = int(hidden_grow_factor * emb_dim)
        if multiple_of:
            hidden = multiple_of * ((hidden + multiple_of - 1) // multiple_of)
        self.hidden_dim = hidden
        self.wg1_fused = nn.Linear(emb_dim, 2 * hidden, bias=False)
        self.w2 = nn.Linear(hidden, emb_dim, bias=False)

    def reset_parameters(self):
        for m in (self.wg1_fused, self.w2):
            nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02)

    def forward(self, x):
        g, u = self.wg1_fused(x).split(self.hidden_dim, dim=-1)
        return self.w2(F.silu(g) * u)
