"""Shard-unit selection ("wrapping") policy.

Reference ``fms_fsdp/policies/wrapping.py:6-14`` hands torch a transformer auto-wrap policy: one FSDP unit per block of
the given class, the remainder (embedding, head, final norm, any other block) in the root unit.  Here the policy is a
predicate that ``ShardedModel(auto_wrap_policy=...)`` evaluates over the model's chain of blocks: accepted blocks become
shard units of their own (gathered just in time, reduced right after their backward), rejected blocks stay resident in
the root unit (``parallel/engine.py``; tested in ``tests/test_engine_cpu.py::test_wrapping_policy_decides_the_shard_units``).
"""
import functools
from typing import Set, Type

import torch.nn as nn


def unit_policy(module: nn.Module, block_classes: Set[Type[nn.Module]]) -> bool:
    return isinstance(module, tuple(block_classes))


def get_wrapper(block):
    """Reference: ``fms_fsdp/policies/wrapping.py:6-14``."""
    return functools.partial(unit_policy, block_classes={block})
