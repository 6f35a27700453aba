"""Export a sharded Mamba training checkpoint in mamba_ssm's ``save_pretrained`` layout
(``config.json`` + ``pytorch_model.bin``); CLI parity with reference ``fms_to_hf_mamba.py:9-37``."""
import torch

from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
from fms_fsdp_b200.utils.cli import run
from fms_fsdp_b200.utils.config_utils import get_model_config
from fms_to_hf_llama import load_dcp_into


def main(model_variant, load_path, save_path, tokenizer_name_or_path=None):
    print("Initializing model...")
    model = MambaLMHeadModel(MambaConfig(**get_model_config(model_variant)))
    print(f"Reading state dict from {load_path}")
    load_dcp_into(model, load_path)
    print("Loading state dict into the model...")
    model.save_pretrained(save_path)
    print(f"Model saving at {save_path}")
    if tokenizer_name_or_path:
        from transformers import AutoTokenizer
        AutoTokenizer.from_pretrained(tokenizer_name_or_path).save_pretrained(save_path)
    print("Done.")


if __name__ == "__main__":
    run(main)
