"""HF export round-trips: logits of the exported HF model == ours (reference leaves this untested; its
exporter drops rope_theta)."""
import os
import tempfile

import pytest
import torch

from fms_fsdp_b200.models.llama import LLaMA, LLaMAConfig
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel
from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer


@pytest.mark.parametrize("kv,theta", [(0, 10000.0), (2, 500000.0)])
def test_llama_export_logits_roundtrip(kv, theta):
    import fms_to_hf_llama as ex
    torch.manual_seed(0)
    cfg = LLaMAConfig(src_vocab_size=97, emb_dim=64, nheads=4, kvheads=kv, nlayers=2, multiple_of=16,
                      max_expected_seq_len=64, rope_theta=theta)
    m = LLaMA(cfg); m.reset_parameters(); m.eval()
    hf = ex.convert_to_hf(m, "llama3_x" if theta > 1e5 else "llama2_x").eval()
    rp = getattr(hf.config, 'rope_parameters', None) or {}
    assert (rp.get('rope_theta') if rp else hf.config.rope_theta) == theta
    x = torch.randint(0, 97, (2, 33))
    with torch.no_grad():
        ours = m(x)
        theirs = hf(x).logits
    assert torch.allclose(ours, theirs, atol=2e-4, rtol=1e-3), (ours - theirs).abs().max()


def test_llama_export_from_sharded_checkpoint(monkeypatch):
    import fms_to_hf_llama as ex
    from fms_fsdp_b200.utils import config_utils
    torch.manual_seed(1)
    cfg = LLaMAConfig(src_vocab_size=64, emb_dim=32, nheads=2, nlayers=2, multiple_of=16, max_expected_seq_len=32)
    m = LLaMA(cfg); m.reset_parameters()
    ref_sd = {k: v.clone() for k, v in m.state_dict().items()}
    eng = ShardedModel(m, device="cpu"); opt = ShardedAdamW(eng)
    ck = tempfile.mkdtemp()
    Checkpointer(ck, 2, "fsdp", 0, 0).save(7, eng, opt, None, tokens_seen=1)
    monkeypatch.setattr(ex, "get_model_config", lambda v: LLaMAConfig(**cfg.__dict__))
    out = tempfile.mkdtemp()
    ex.main("llama2_test", load_path=os.path.join(ck, "checkpoints", "step_7_ckp"), save_path=out)
    from transformers import LlamaForCausalLM
    hf = LlamaForCausalLM.from_pretrained(out)
    assert torch.equal(hf.model.embed_tokens.weight, ref_sd["shared.emb.weight"])
    assert torch.equal(hf.lm_head.weight, ref_sd["shared.head.weight"])


def test_mamba_export(monkeypatch):
    import fms_to_hf_mamba as ex
    from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
    from fms_fsdp_b200.utils.config_utils import get_model_config
    torch.manual_seed(2)
    m = MambaLMHeadModel(MambaConfig(**get_model_config("mamba_tiny"))); m.reset_parameters()
    ref_sd = {k: v.clone() for k, v in m.state_dict().items()}
    eng = ShardedModel(m, device="cpu")
    ck = tempfile.mkdtemp()
    Checkpointer(ck, 2, "fsdp", 0, 0).save(3, eng, None, None)
    out = tempfile.mkdtemp()
    ex.main("mamba_tiny", os.path.join(ck, "checkpoints", "step_3_ckp"), out)
    sd = torch.load(os.path.join(out, "pytorch_model.bin"))
    assert set(sd) == set(ref_sd) and all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    assert os.path.exists(os.path.join(out, "config.json"))


def test_hf_import_single_file_feeds_the_trainer(tmp_path):
    """HF Llama directory -> ``hf_to_fms_llama.py`` -> single-file checkpoint -> ``Checkpointer.load`` into the sharded runtime:
    the engine's logits equal the HF model's, and the load is reported as a fresh start (not a resume)."""
    import hf_to_fms_llama as imp
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(3)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=80, hidden_size=64, intermediate_size=96, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                                      rope_theta=50000.0, tie_word_embeddings=False)).eval()
    hf_dir, pth = str(tmp_path / "hf"), str(tmp_path / "w" / "llama.pth")
    hf.save_pretrained(hf_dir)
    imp.main(hf_dir, pth, dtype="fp32")

    from fms_fsdp_b200.models.hf_loader import config_from_hf
    cfg = config_from_hf(hf.config.to_dict())
    m = LLaMA(cfg); m.reset_parameters()
    eng = ShardedModel(m, device="cpu"); opt = ShardedAdamW(eng)
    _, _, _, step, tokens, resuming = Checkpointer(str(tmp_path / "save"), 2, "fsdp", 0, 0).load(eng, opt, None, path=pth)
    assert (step, tokens, resuming) == (0, 0, False)
    x = torch.randint(0, 80, (2, 21))
    with torch.no_grad():
        theirs = hf(x).logits
        ours = eng.module(x)
    assert torch.allclose(ours.float(), theirs, atol=2e-4, rtol=1e-3), (ours - theirs).abs().max()

    with pytest.raises(ValueError, match="does not have the architecture"):
        imp.main(hf_dir, pth, model_variant="llama2_7b")


@pytest.mark.parametrize("compiled,old_fms", [(True, False), (False, True), (True, True)])
def test_llama_export_reads_compiled_and_pre_fusion_checkpoints(monkeypatch, compiled, old_fms):
    """``--compiled`` (keys under ``model_state._orig_mod.``) and ``--is_old_fms`` (separate q/k/v and gate/up matrices on disk,
    reference ``fms_to_hf_llama.py:60-95``) checkpoints export to the same HF model as the plain fused checkpoint."""
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint import FileSystemWriter
    import fms_to_hf_llama as ex
    torch.manual_seed(4)
    cfg = LLaMAConfig(src_vocab_size=64, emb_dim=32, nheads=4, kvheads=2, nlayers=2, multiple_of=16, max_expected_seq_len=32)
    m = LLaMA(cfg); m.reset_parameters(); m.eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    if old_fms:
        hd, split = cfg.head_dim, {}
        for k, v in sd.items():
            if k.endswith("in_proj.qkv_fused.weight"):
                q, kk, vv = torch.split(v, [cfg.nheads * hd, cfg.kv_heads * hd, cfg.kv_heads * hd])
                base = k[:-len("in_proj.qkv_fused.weight")]
                split.update({base + "query.weight": q.clone(), base + "key.weight": kk.clone(), base + "value.weight": vv.clone()})
            elif k.endswith("wg1_fused.weight"):
                g, u = v.chunk(2)
                split.update({k[:-len("wg1_fused.weight")] + "wg.weight": g.clone(), k[:-len("wg1_fused.weight")] + "w1.weight": u.clone()})
            else:
                split[k] = v
        sd = split
    ck = tempfile.mkdtemp()
    dcp.save({"model_state": {"_orig_mod": sd} if compiled else sd},
             storage_writer=FileSystemWriter(ck, single_file_per_rank=True), no_dist=True)
    monkeypatch.setattr(ex, "get_model_config", lambda v: LLaMAConfig(**cfg.__dict__))
    out = tempfile.mkdtemp()
    ex.main("llama2_test", compiled, old_fms, ck, out)          # the reference's positional order
    from transformers import LlamaForCausalLM
    hf = LlamaForCausalLM.from_pretrained(out).eval()
    x = torch.randint(0, 64, (2, 17))
    with torch.no_grad():
        assert torch.allclose(m(x), hf(x).logits, atol=2e-4, rtol=1e-3)


def _init(m):
    m.reset_parameters()
    return m


def test_mamba_export_in_transformers_format(monkeypatch):
    """``fms_to_hf_mamba.py --transformers_format``: a pure-Mamba2 checkpoint becomes a directory
    ``transformers.Mamba2ForCausalLM.from_pretrained`` loads, with the same logits; hybrid models are refused with a reason."""
    import fms_to_hf_mamba as ex
    from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
    from fms_fsdp_b200.utils.config_utils import get_model_config
    pure = dict(d_model=64, d_intermediate=0, n_layer=2, vocab_size=256, attn_layer_idx=[], attn_cfg={},
                ssm_cfg={"layer": "Mamba2", "headdim": 16, "d_state": 16, "chunk_size": 16}, rms_norm=True, residual_in_fp32=True,
                fused_add_norm=True, pad_vocab_size_multiple=16, tie_embeddings=True)
    torch.manual_seed(5)
    m = MambaLMHeadModel(MambaConfig(**pure)); m.reset_parameters(); m.eval()
    eng = ShardedModel(m, device="cpu")
    ck = tempfile.mkdtemp()
    Checkpointer(ck, 2, "fsdp", 0, 0).save(1, eng, None, None)
    real = get_model_config
    monkeypatch.setattr(ex, "get_model_config", lambda v: dict(pure) if v == "mamba_pure_test" else real(v))
    out = tempfile.mkdtemp()
    ex.main("mamba_pure_test", os.path.join(ck, "checkpoints", "step_1_ckp"), out, transformers_format=True)
    from transformers import Mamba2ForCausalLM
    hf = Mamba2ForCausalLM.from_pretrained(out).eval()
    x = torch.randint(0, 256, (2, 23))
    with torch.no_grad():
        with eng.summon_full_params():
            a = eng.module(x)
        a = a.logits if hasattr(a, "logits") else a
        b = hf(x).logits
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-4), (a - b).abs().max()
    assert type(ex.to_transformers(_init(MambaLMHeadModel(MambaConfig(**get_model_config("mamba_tiny")))))).__name__ == "BambaForCausalLM"
    with pytest.raises(ValueError, match="pure Mamba1 / Mamba2 stack"):      # attention layers but no MLP: neither layout fits
        ex.to_transformers(MambaLMHeadModel(MambaConfig(**{**get_model_config("mamba_tiny"), "d_intermediate": 0})))
    # the tied head is stored once; the mamba_ssm-layout export and a trainer resume both restore the alias
    out2 = tempfile.mkdtemp()
    ex.main("mamba_pure_test", os.path.join(ck, "checkpoints", "step_1_ckp"), out2)
    sd = torch.load(os.path.join(out2, "pytorch_model.bin"))
    assert torch.equal(sd["lm_head.weight"], sd["backbone.embedding.weight"])
    torch.manual_seed(9)
    m2 = MambaLMHeadModel(MambaConfig(**pure)); m2.reset_parameters()
    eng2 = ShardedModel(m2, device="cpu")
    Checkpointer(tempfile.mkdtemp(), 2, "fsdp", 0, 0).load(eng2, None, None, path=os.path.join(ck, "checkpoints", "step_1_ckp"))
    a2, b2 = eng.full_state_dict(), eng2.full_state_dict()
    assert set(a2) == set(b2) and all(torch.equal(a2[k], b2[k]) for k in a2)


@pytest.mark.parametrize("scaling", [
    {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0, "original_max_position_embeddings": 32},
    {"rope_type": "linear", "factor": 4.0}])
def test_rope_scaled_hf_checkpoints_load_and_export_with_equal_logits(tmp_path, scaling):
    """Llama 3.1-style (``llama3``) and ``linear`` rope scaling: an HF checkpoint that uses it loads with the same logits
    (the scaling only changes the cos / sin table the kernels read), and exporting it again reproduces the HF model."""
    import fms_to_hf_llama as ex
    from transformers import LlamaConfig, LlamaForCausalLM
    from fms_fsdp_b200.models.hf_loader import load_hf_llama
    torch.manual_seed(6)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=96, hidden_size=64, intermediate_size=96, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
                                      rope_theta=10000.0, rope_scaling=dict(scaling), tie_word_embeddings=False)).eval()
    d = str(tmp_path / "hf")
    hf.save_pretrained(d)
    ours = load_hf_llama(d, "cpu", torch.float32).eval()
    assert ours.config.rope_scaling and (ours.config.rope_scaling.get("rope_type") == scaling["rope_type"])
    x = torch.randint(0, 96, (2, 200))                      # long enough that the rescaled low frequencies matter
    with torch.no_grad():
        want = hf(x).logits
        got = ours(x)
        plain = LLaMA(LLaMAConfig(**{**ours.config.__dict__, "rope_scaling": None}))
        plain.load_state_dict(ours.state_dict())
        unscaled = plain.eval()(x)
    assert torch.allclose(got, want, atol=3e-4, rtol=1e-3), (got - want).abs().max()
    assert (unscaled - want).abs().max() > 10 * (got - want).abs().max()        # the scaling is not a no-op here
    back = ex.convert_to_hf(ours, "llama3_x").eval()
    with torch.no_grad():
        assert torch.allclose(back(x).logits, want, atol=3e-4, rtol=1e-3)
