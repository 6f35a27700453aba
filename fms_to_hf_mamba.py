"""Export a sharded Mamba training checkpoint in mamba_ssm's ``save_pretrained`` layout
(``config.json`` + ``pytorch_model.bin``); CLI parity with reference ``fms_to_hf_mamba.py:9-37``.

``--transformers_format`` (extension) writes a directory ``transformers.Mamba2ForCausalLM.from_pretrained`` loads instead --
possible for pure Mamba2 stacks (no attention layers, no MLP: e.g. ``mamba_2.8b``); the parameter names coincide except for
the embedding, and the logits of the two implementations agree (``tests/test_mamba.py``)."""
import torch

from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
from fms_fsdp_b200.utils.cli import run
from fms_fsdp_b200.utils.config_utils import get_model_config
from fms_to_hf_llama import load_dcp_into


def to_transformers(model: MambaLMHeadModel):
    """``transformers.Mamba2ForCausalLM`` carrying the weights of a pure-Mamba2 ``MambaLMHeadModel``."""
    from transformers import Mamba2Config, Mamba2ForCausalLM
    c = model.config
    if c.attn_layer_idx or c.d_intermediate or (c.ssm_cfg or {}).get("layer", "Mamba2") != "Mamba2":
        raise ValueError("--transformers_format needs a pure Mamba2 stack (no attention layers, d_intermediate = 0); "
                         "hybrid models export in the mamba_ssm layout")
    mixer = model.backbone.layers[0].mixer
    hf = Mamba2ForCausalLM(Mamba2Config(
        vocab_size=c.padded_vocab, hidden_size=c.d_model, state_size=mixer.d_state, num_hidden_layers=c.n_layer,
        head_dim=mixer.headdim, num_heads=mixer.nheads, expand=mixer.expand, n_groups=mixer.ngroups, conv_kernel=mixer.d_conv,
        chunk_size=mixer.chunk_size, tie_word_embeddings=bool(c.tie_embeddings), rms_norm=bool(c.rms_norm), use_bias=False,
        use_conv_bias=mixer.conv1d.bias is not None, residual_in_fp32=bool(c.residual_in_fp32),
        layer_norm_epsilon=c.norm_epsilon, hidden_act="silu"))
    sd = {("backbone.embeddings.weight" if k == "backbone.embedding.weight" else k): v for k, v in model.state_dict().items()}
    hf.load_state_dict(sd, strict=not c.tie_embeddings)
    return hf


def main(model_variant, load_path, save_path, tokenizer_name_or_path=None, transformers_format=False):
    print("Initializing model...")
    model = MambaLMHeadModel(MambaConfig(**get_model_config(model_variant)))
    print(f"Reading state dict from {load_path}")
    load_dcp_into(model, load_path)
    print("Loading state dict into the model...")
    if transformers_format:
        to_transformers(model).save_pretrained(save_path)
    else:
        model.save_pretrained(save_path)
    print(f"Model saving at {save_path}")
    if tokenizer_name_or_path:
        from transformers import AutoTokenizer
        AutoTokenizer.from_pretrained(tokenizer_name_or_path).save_pretrained(save_path)
    print("Done.")


if __name__ == "__main__":
    run(main)
