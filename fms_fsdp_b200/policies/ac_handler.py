"""Selective activation checkpointing (recompute) placement.

Same selection rule as reference ``fms_fsdp/policies/ac_handler.py:43-58`` -- evenly spaced
fraction ``p`` of the blocks, patterns pinned by ``tests/test_selective_ac.py`` -- but the
mechanism is the engine's own: a selected block is flagged, the runtime then keeps only the
block input in forward and re-runs the fused block forward inside backward *while the
block's gathered weights are already resident for the backward pass* (no second all-gather).
"""
from fractions import Fraction
from typing import Union

import torch.nn as nn

_FLAG = "_b200_recompute"


def parse_fraction(p: Union[int, float, str]) -> float:
    """'1/3' -> 0.333..; numbers pass through. (The reference eval()s the CLI string.)"""
    if isinstance(p, str):
        return float(Fraction(p.strip()))
    return float(p)


def selection_mask(n_blocks: int, p: Union[int, float, str]):
    """Boolean list: block i (1-based counter) is selected iff i*p >= cut_off, cut_off starting at
    1/2 and advancing by 1 after each selection."""
    p = parse_fraction(p)
    mask, cut_off = [], 0.5
    for i in range(1, n_blocks + 1):
        if i * p >= cut_off:
            cut_off += 1
            mask.append(True)
        else:
            mask.append(False)
    return mask


def non_reentrant_wrapper(module: nn.Module) -> nn.Module:
    """Counterpart of the reference's module-level ``non_reentrant_wrapper`` (``ac_handler.py:10-13``, a partial of
    torch's ``checkpoint_wrapper``): marks ONE block for recompute and returns it (no wrapper module is inserted, so
    state-dict keys do not change)."""
    setattr(module, _FLAG, True)
    return module


def is_checkpointed(module: nn.Module) -> bool:
    return bool(getattr(module, _FLAG, False))


def apply_fsdp_checkpointing(model: nn.Module, block, p):
    """Flag fraction ``p`` of ``block`` instances under ``model`` for recompute.
    Reference: ``fms_fsdp/policies/ac_handler.py:16-64``."""
    blocks = [m for m in model.modules() if isinstance(m, block)]
    for m, sel in zip(blocks, selection_mask(len(blocks), p)):
        setattr(m, _FLAG, bool(sel))
    return model
