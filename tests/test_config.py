import pytest

from fms_fsdp_b200.config import train_config
from fms_fsdp_b200.utils.cli import parse_argv
from fms_fsdp_b200.utils.config_utils import get_model_config, list_model_variants, update_config


def test_defaults_match_reference_surface():
    c = train_config()
    assert (c.seq_length, c.batch_size, c.sharding_strategy, c.learning_rate) == (4096, 2, "hsdp", 3e-4)
    assert c.selective_checkpointing == 1 and c.logical_shards == 1024 and c.eos_token == 0
    assert len(c.datasets.split(",")) == 13 == len(c.weights.split(","))


def test_update_config_known_unknown_dotted(capsys):
    c = train_config()
    update_config(c, batch_size=4, **{"train_config.seq_length": 128, "bogus": 1})
    assert c.batch_size == 4 and c.seq_length == 128
    assert "unknown parameter bogus" in capsys.readouterr().out


def test_cli_literal_typing():
    kw = parse_argv(["--batch_size=4", "--learning_rate", "3e-4", "--use_dummy_dataset", "--nolow_cpu_fsdp",
                     "--selective_checkpointing=1/3", "--tracker=None", "--model_variant=llama2_7b"])
    assert kw == dict(batch_size=4, learning_rate=3e-4, use_dummy_dataset=True, low_cpu_fsdp=False,
                      selective_checkpointing="1/3", tracker=None, model_variant="llama2_7b")


def test_model_zoo_shapes():
    c = get_model_config("llama2_7b")
    assert (c.emb_dim, c.nheads, c.kv_heads, c.nlayers, c.hidden_dim) == (4096, 32, 32, 32, 11008)
    c = get_model_config("llama2_70b")
    assert (c.hidden_dim, c.kv_heads) == (28672, 8)
    c = get_model_config("llama3_8b")
    assert (c.hidden_dim, c.src_vocab_size, c.rope_theta) == (14336, 128256, 500000.0)
    assert get_model_config("mamba_9.8b")["attn_layer_idx"] == [9, 18, 27]
    assert "mamba_2.8b" in list_model_variants() and "llama2_tiny" in list_model_variants()
    with pytest.raises(ValueError):
        get_model_config("7b")  # the reference's own default is not a valid key


def test_param_counts():
    import torch
    from fms_fsdp_b200.models.llama import LLaMA
    with torch.device("meta"):
        m = LLaMA(get_model_config("llama2_7b"))
    assert sum(p.numel() for p in m.parameters()) == 6738415616


def test_grad_dtype_and_precision_switches():
    import torch
    from fms_fsdp_b200.config import train_config
    from fms_fsdp_b200.models.llama import LLaMABlock
    from fms_fsdp_b200.utils.train_utils import get_policies
    cfg = train_config()
    assert get_policies(cfg, 1, LLaMABlock)[0].reduce_dtype == torch.bfloat16
    cfg.grad_dtype = "fp32"
    mp = get_policies(cfg, 1, LLaMABlock)[0]
    assert mp.reduce_dtype == torch.float32 and mp.param_dtype == torch.bfloat16
    cfg.precision = "mxfp8"
    import pytest
    with pytest.raises(NotImplementedError):
        get_policies(cfg, 1, LLaMABlock)
