from fms_fsdp_b200.config.training import train_config

__all__ = ["train_config"]
