import torch
import torch.nn as nn


class LayerNormParameterized(nn.Module):
    """RMSNorm flavour used by fms LLaMA (use_mean=False, elementwise_scale=True, no shift); fp32 math."""

    def __init__(self, normalized_shape, eps=1e-6, elementwise_scale=True, elementwise_shift=False, use_mean=False,
                 use_high_precision_pow=True):
        super().__init__()
        self.normalized_shape, self.eps = normalized_shape, eps
        self.weight = nn.Parameter(torch.empty(normalized_shape))

    def reset_parameters(self):
        self.weight.data.fill_(1)

    def forward(self, x):
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        return self.weight * xf.type_as(x)
