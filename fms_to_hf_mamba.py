"""Export a sharded Mamba training checkpoint in mamba_ssm's ``save_pretrained`` layout
(``config.json`` + ``pytorch_model.bin``); CLI parity with reference ``fms_to_hf_mamba.py:9-37``.

``--transformers_format`` (extension) writes a directory ``transformers`` loads instead: ``Mamba2ForCausalLM`` for pure Mamba2
stacks (``mamba_2.8b``), ``BambaForCausalLM`` for the Mamba2 + attention + MLP hybrid (``mamba_9.8b``).  The logits of this
repo's model and of the ``transformers`` implementations agree with the same weights (``tests/test_mamba.py``)."""
import torch

from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
from fms_fsdp_b200.utils.cli import run
from fms_fsdp_b200.utils.config_utils import get_model_config
from fms_to_hf_llama import load_dcp_into


def _to_mamba2(model: MambaLMHeadModel):
    from transformers import Mamba2Config, Mamba2ForCausalLM
    c = model.config
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


def _to_mamba1(model: MambaLMHeadModel):
    from transformers import MambaConfig as HFMambaConfig
    from transformers import MambaForCausalLM
    c = model.config
    mixer = model.backbone.layers[0].mixer
    hf = MambaForCausalLM(HFMambaConfig(
        vocab_size=c.padded_vocab, hidden_size=c.d_model, state_size=mixer.d_state, num_hidden_layers=c.n_layer,
        expand=mixer.expand, conv_kernel=mixer.d_conv, use_bias=False, use_conv_bias=mixer.conv1d.bias is not None,
        time_step_rank=mixer.dt_rank, residual_in_fp32=bool(c.residual_in_fp32), layer_norm_epsilon=c.norm_epsilon,
        hidden_act="silu", tie_word_embeddings=bool(c.tie_embeddings), use_mambapy=False))
    sd = {("backbone.embeddings.weight" if k == "backbone.embedding.weight" else k): v for k, v in model.state_dict().items()}
    hf.load_state_dict(sd, strict=not c.tie_embeddings)
    return hf


def _to_bamba(model: MambaLMHeadModel):
    """Mamba2 + attention + gated-MLP hybrid (``mamba_9.8b``) -> ``transformers.BambaForCausalLM``: fused attention ``in_proj``
    is split into q / k / v, ``mlp.fc1`` = (value | gate) halves become ``up_proj`` / ``gate_proj``, the two block norms become
    ``input_layernorm`` / ``pre_ff_layernorm``."""
    import re

    from transformers import BambaConfig, BambaForCausalLM
    c = model.config
    a = dict(c.attn_cfg)
    if not c.d_intermediate or a.get("d_conv", 0) or a.get("qkv_proj_bias") or a.get("out_proj_bias"):
        raise ValueError("the Bamba layout needs an MLP in every block and bias-free, conv-free attention layers")
    mamba = next(b.mixer for i, b in enumerate(model.backbone.layers) if i not in c.attn_layer_idx)
    H, KV = a["num_heads"], a.get("num_heads_kv") or a["num_heads"]
    hd = a.get("head_dim") or c.d_model // H
    hf = BambaForCausalLM(BambaConfig(
        vocab_size=c.padded_vocab, hidden_size=c.d_model, intermediate_size=c.d_intermediate, num_hidden_layers=c.n_layer,
        num_attention_heads=H, num_key_value_heads=KV, attn_layer_indices=list(c.attn_layer_idx), mamba_n_heads=mamba.nheads,
        mamba_d_head=mamba.headdim, mamba_n_groups=mamba.ngroups, mamba_d_state=mamba.d_state, mamba_d_conv=mamba.d_conv,
        mamba_expand=mamba.expand, mamba_chunk_size=mamba.chunk_size, mamba_conv_bias=mamba.conv1d.bias is not None,
        mamba_proj_bias=False, tie_word_embeddings=bool(c.tie_embeddings), rms_norm_eps=c.norm_epsilon, attention_bias=False,
        mlp_bias=False, attention_dropout=0.0, partial_rotary_factor=a.get("rotary_emb_dim", 0) / hd,
        rope_theta=float(a.get("rotary_emb_base", 10000.0))))
    sd = {}
    for k, v in model.state_dict().items():
        k = k.replace("backbone.embedding.", "model.embed_tokens.").replace("backbone.norm_f.", "model.final_layernorm.")
        m = re.match(r"backbone\.layers\.(\d+)\.(.*)", k)
        if not m:
            sd[k] = v
            continue
        i, rest = int(m.group(1)), m.group(2)
        pre = f"model.layers.{i}."
        if rest == "norm.weight":
            sd[pre + "input_layernorm.weight"] = v
        elif rest == "norm2.weight":
            sd[pre + "pre_ff_layernorm.weight"] = v
        elif rest == "mlp.fc1.weight":
            sd[pre + "feed_forward.up_proj.weight"], sd[pre + "feed_forward.gate_proj.weight"] = v.chunk(2, dim=0)
        elif rest == "mlp.fc2.weight":
            sd[pre + "feed_forward.down_proj.weight"] = v
        elif i in c.attn_layer_idx and rest == "mixer.in_proj.weight":
            q, kk, vv = v.split([H * hd, KV * hd, KV * hd], dim=0)
            sd[pre + "self_attn.q_proj.weight"], sd[pre + "self_attn.k_proj.weight"], sd[pre + "self_attn.v_proj.weight"] = q, kk, vv
        elif i in c.attn_layer_idx and rest == "mixer.out_proj.weight":
            sd[pre + "self_attn.o_proj.weight"] = v
        else:
            sd[pre + rest.replace("mixer.", "mamba.", 1)] = v
    hf.load_state_dict(sd, strict=not c.tie_embeddings)
    return hf


def to_transformers(model: MambaLMHeadModel):
    """A ``transformers`` model carrying the weights: ``MambaForCausalLM`` / ``Mamba2ForCausalLM`` for pure Mamba1 / Mamba2
    stacks, ``BambaForCausalLM`` for the Mamba2 + attention + MLP hybrid.  Same logits as this repo's model
    (``tests/test_mamba.py``).  For the hybrid that holds with full rotary: transformers 5.5 builds full-width rotary tables
    for Bamba whatever ``partial_rotary_factor`` says, so a partially rotated export reproduces there only what that version
    computes."""
    c = model.config
    kind = (c.ssm_cfg or {}).get("layer", "Mamba1")
    if not c.attn_layer_idx and not c.d_intermediate:
        return _to_mamba2(model) if kind == "Mamba2" else _to_mamba1(model)
    if kind != "Mamba2":
        raise ValueError("--transformers_format: hybrids are exported as Bamba, which needs Mamba2 mixers")
    if c.attn_layer_idx and c.d_intermediate:
        return _to_bamba(model)
    raise ValueError("--transformers_format covers a pure Mamba1 / Mamba2 stack (no attention layers, d_intermediate = 0) or the full "
                     "hybrid (attention layers + MLP in every block); other mixes export in the mamba_ssm layout")


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
