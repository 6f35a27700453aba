"""Load a HuggingFace Llama checkpoint (safetensors / .bin) into the engine's LLaMA (stand-in for
``fms.models.get_model(..., source="hf")``): fuses q/k/v and gate/up, and permutes
q/k rows from HF's half-split RoPE layout back to the FMS interleaved-pair layout (the inverse of
``fms_to_hf_llama._interleaved_to_halfsplit``)."""
from __future__ import annotations

import glob
import json
import os
from typing import Dict

import torch

from fms_fsdp_b200.models.llama import LLaMA, LLaMAConfig


def _halfsplit_to_interleaved(w: torch.Tensor, nheads: int) -> torch.Tensor:
    return w.view(nheads, 2, -1, w.size(1)).transpose(1, 2).reshape(*w.size())


def config_from_hf(hf_cfg: dict) -> LLaMAConfig:
    D, F = hf_cfg["hidden_size"], hf_cfg["intermediate_size"]
    rope = hf_cfg.get("rope_theta") or (hf_cfg.get("rope_parameters") or {}).get("rope_theta", 10000.0)
    return LLaMAConfig(
        src_vocab_size=hf_cfg["vocab_size"], emb_dim=D, norm_eps=hf_cfg.get("rms_norm_eps", 1e-5),
        nheads=hf_cfg["num_attention_heads"], kvheads=hf_cfg.get("num_key_value_heads", 0) or 0,
        nlayers=hf_cfg["num_hidden_layers"], hidden_grow_factor=F / D, multiple_of=1,
        max_expected_seq_len=hf_cfg.get("max_position_embeddings", 4096), rope_theta=float(rope))


def _read_hf_tensors(model_path: str) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    st = sorted(glob.glob(os.path.join(model_path, "*.safetensors")))
    if st:
        from safetensors.torch import load_file
        for f in st:
            out.update(load_file(f))
        return out
    for f in sorted(glob.glob(os.path.join(model_path, "pytorch_model*.bin"))):
        out.update(torch.load(f, map_location="cpu", weights_only=True))
    if not out:
        raise FileNotFoundError(f"no safetensors / pytorch_model*.bin under {model_path}")
    return out


def convert_hf_state_dict(hf: Dict[str, torch.Tensor], cfg: LLaMAConfig) -> Dict[str, torch.Tensor]:
    sd = {"shared.emb.weight": hf["model.embed_tokens.weight"],
          "shared.head.weight": hf.get("lm_head.weight", hf["model.embed_tokens.weight"]),
          "dec_norm.weight": hf["model.norm.weight"]}
    for i in range(cfg.nlayers):
        s, d = f"model.layers.{i}.", f"layers.{i}."
        q = _halfsplit_to_interleaved(hf[s + "self_attn.q_proj.weight"], cfg.nheads)
        k = _halfsplit_to_interleaved(hf[s + "self_attn.k_proj.weight"], cfg.kv_heads)
        sd[d + "attn.in_proj.qkv_fused.weight"] = torch.cat([q, k, hf[s + "self_attn.v_proj.weight"]], dim=0)
        sd[d + "attn.dense.weight"] = hf[s + "self_attn.o_proj.weight"]
        sd[d + "ff_sub_layer.wg1_fused.weight"] = torch.cat([hf[s + "mlp.gate_proj.weight"], hf[s + "mlp.up_proj.weight"]], dim=0)
        sd[d + "ff_sub_layer.w2.weight"] = hf[s + "mlp.down_proj.weight"]
        sd[d + "ln.weight"] = hf[s + "input_layernorm.weight"]
        sd[d + "ff_ln.weight"] = hf[s + "post_attention_layernorm.weight"]
    return sd


def load_hf_llama(model_path: str, device="cpu", dtype=torch.bfloat16, model_cls=LLaMA) -> LLaMA:
    with open(os.path.join(model_path, "config.json")) as f:
        cfg = config_from_hf(json.load(f))
    with torch.device("meta"):
        model = model_cls(cfg)
    model.to_empty(device=device)
    sd = convert_hf_state_dict(_read_hf_tensors(model_path), cfg)
    model.load_state_dict({k: v.to(dtype) for k, v in sd.items()})
    return model.to(dtype)
