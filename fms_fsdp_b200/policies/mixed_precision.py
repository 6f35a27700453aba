"""Mixed-precision policies (names from reference ``fms_fsdp/policies/mixed_precision.py:5-27``).

``param_dtype``  dtype of the gathered parameters / compute (the low-precision shard each rank keeps),
``reduce_dtype`` dtype of the unsharded gradient buffer that enters the reduce-scatter,
``buffer_dtype`` dtype of non-parameter buffers.  Master weights and AdamW moments are always fp32.
"""
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class MixedPrecision:
    param_dtype: torch.dtype = torch.float32
    reduce_dtype: torch.dtype = torch.float32
    buffer_dtype: torch.dtype = torch.float32


fpSixteen = MixedPrecision(torch.float16, torch.float16, torch.float16)
bfSixteen = MixedPrecision(torch.bfloat16, torch.bfloat16, torch.bfloat16)
bfSixteen_working = MixedPrecision(torch.float32, torch.bfloat16, torch.bfloat16)
fp32_policy = MixedPrecision(torch.float32, torch.float32, torch.float32)
