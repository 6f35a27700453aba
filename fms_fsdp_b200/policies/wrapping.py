"""Shard-unit selection ("wrapping") policy.

Reference ``fms_fsdp/policies/wrapping.py:6-14`` hands torch a transformer auto-wrap policy: one
FSDP unit per block, remainder in the root.  Here the policy is a predicate the sharded runtime
evaluates to pick unit boundaries.
"""
import functools
from typing import Set, Type

import torch.nn as nn


def unit_policy(module: nn.Module, block_classes: Set[Type[nn.Module]]) -> bool:
    return isinstance(module, tuple(block_classes))


def get_wrapper(block):
    return functools.partial(unit_policy, block_classes={block})
