"""Deferred (meta-device) parameter materialisation.

Reference ``fms_fsdp/policies/param_init.py:9-18``: modules built on the meta device are
``to_empty``'d on the GPU and ``reset_parameters()``'d when ``low_cpu_fsdp`` is set.  The
sharded runtime calls this per shard unit, so at most one unit is ever materialised unsharded.
"""
import torch
import torch.nn as nn


def param_init_function(module: nn.Module, device=None):
    """Reference: ``fms_fsdp/policies/param_init.py:9-18``."""
    device = device if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    if any(p.is_meta for p in module.parameters(recurse=False)) or any(
            b.is_meta for b in module.buffers(recurse=False)):
        module.to_empty(device=device, recurse=False)
    if hasattr(module, "reset_parameters") and len(list(module.parameters(recurse=False))) + len(
            list(module.children())) > 0:
        try:
            module.reset_parameters()
        except NotImplementedError:
            pass
