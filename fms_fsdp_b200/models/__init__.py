from fms_fsdp_b200.models.llama import LLaMA, LLaMABlock, LLaMAConfig  # noqa: F401
