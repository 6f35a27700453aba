from fms_fsdp_b200.policies.ac_handler import apply_fsdp_checkpointing, is_checkpointed, selection_mask
from fms_fsdp_b200.policies.mixed_precision import *  # noqa: F401,F403  (reference star-exports these)
from fms_fsdp_b200.policies.mixed_precision import MixedPrecision, bfSixteen, bfSixteen_working, fp32_policy, fpSixteen
from fms_fsdp_b200.policies.param_init import param_init_function
from fms_fsdp_b200.policies.wrapping import get_wrapper
