"""Speculator stack on a tiny frozen base model: both training stages learn, KV-cache generation matches
full re-forward, HF load inverts the export permutation."""
import os
import tempfile

import pytest
import torch

from fms_fsdp_b200.config import train_config
from fms_fsdp_b200.models.llama import LLaMAConfig
from fms_fsdp_b200.models.speculator import MLPSpeculator
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel
from speculator.train_speculator import speculator_lr_schedule
from speculator.train_speculator_utils import (EmbedGPTBigCode, EmbedLLaMA, EmbedMixtral, do_ckpt, generate, stage1_loss,
                                               stage2_loss, train_speculator)


def _base():
    torch.manual_seed(0)
    m = EmbedLLaMA(LLaMAConfig(src_vocab_size=64, emb_dim=32, nheads=4, kvheads=2, nlayers=2, multiple_of=8,
                               max_expected_seq_len=128))
    m.reset_parameters()
    for p in m.parameters():
        p.requires_grad_(False)
    return m.eval()


def test_kv_cache_generation_matches_full_forward():
    m = _base()
    prompt = torch.randint(0, 64, (2, 7))
    a, ea = generate(m, prompt, max_new_tokens=9, do_sample=False, use_cache=True, include_embeds=True)
    b, eb = generate(m, prompt, max_new_tokens=9, do_sample=False, use_cache=False, include_embeds=True)
    assert torch.equal(a, b)
    # cached path returns one embedding per step; uncached returns whole-sequence embeds each step
    logits, emb = m(a[:, :-1], include_embeds=True)
    assert torch.allclose(ea[:, -1], emb[:, -1], atol=1e-4)


@pytest.mark.parametrize("stage", [1, 2])
def test_stage_losses_decrease(stage):
    base = _base()
    cfg = train_config()
    cfg.n_speculator_heads, cfg.batch_size, cfg.seq_length = 2, 2, 24 + 3
    cfg.stage2_batch_size, cfg.stage2_prompt_length, cfg.stage2_seq_length = 4, 6, 10
    cfg.sharding_strategy = "ddp"
    spec = MLPSpeculator(32, 24, 64, 2, tie_weights=True, scale_input=True); spec.reset_parameters()
    eng = ShardedModel(spec, sharding_strategy="ddp", device="cpu"); opt = ShardedAdamW(eng, lr=2e-2)
    torch.manual_seed(1)
    x = torch.randint(0, 64, (2, cfg.seq_length))
    stats = torch.zeros(4)
    fn = stage1_loss if stage == 1 else stage2_loss
    losses = []
    for _ in range(12):
        torch.manual_seed(5)  # same sampled continuation every step in stage 2
        l = eng.forward_backward_custom(lambda mod: fn(cfg, base, mod, x, x, torch.nn.CrossEntropyLoss(), stats, None)[0])
        eng.clip_grad_norm_(1.0); opt.step(); losses.append(l.item())
    assert losses[-1] < 0.8 * losses[0]


def test_train_speculator_loop_and_checkpoint(capsys):
    base = _base()
    cfg = train_config()
    cfg.n_speculator_heads, cfg.batch_size, cfg.seq_length, cfg.vocab_size = 2, 2, 16 + 3, 64
    cfg.num_steps, cfg.report_interval, cfg.checkpoint_interval, cfg.stage2_start_step = 6, 3, 100, 4
    cfg.stage2_batch_size, cfg.stage2_prompt_length, cfg.stage2_seq_length = 4, 4, 8
    cfg.sharding_strategy = "ddp"
    cfg.ckpt_save_path = tempfile.mkdtemp()
    spec = MLPSpeculator(32, 24, 64, 2); spec.reset_parameters()
    eng = ShardedModel(spec, sharding_strategy="ddp", device="cpu"); opt = ShardedAdamW(eng, lr=1e-2)
    from torch.optim.lr_scheduler import LambdaLR
    from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
    sched = LambdaLR(opt, speculator_lr_schedule(cfg))
    loader = (torch.randint(0, 64, (2, cfg.seq_length)) for _ in range(100))
    train_speculator(cfg, base, eng, 0, 0, loader, opt, sched, Checkpointer(cfg.ckpt_save_path, 3, "ddp", 0, 0))
    out = capsys.readouterr().out
    assert "loss 1:" in out and "loss 2:" in out and "step: 6" in out
    assert os.path.isdir(os.path.join(cfg.ckpt_save_path, "checkpoints", "step_6_ckp"))
    open(cfg.ckpt_save_path + "/do_ckpt", "w").write("1")
    assert do_ckpt(cfg.ckpt_save_path) and not do_ckpt(cfg.ckpt_save_path, reset=True) and not do_ckpt(cfg.ckpt_save_path)


def test_other_registered_archs_forward():
    g = EmbedGPTBigCode(vocab=50, emb_dim=32, nheads=4, nlayers=2, max_pos=64)
    x = torch.randint(0, 50, (2, 9))
    lg, emb = g(x, include_embeds=True)
    assert lg.shape == (2, 9, 50) and emb.shape == (2, 9, 32)
    a, _ = generate(g, x, max_new_tokens=3, do_sample=False, use_cache=True)
    b, _ = generate(g, x, max_new_tokens=3, do_sample=False, use_cache=False)
    assert torch.equal(a, b)
    mx = EmbedMixtral(LLaMAConfig(src_vocab_size=50, emb_dim=32, nheads=4, kvheads=2, nlayers=2, multiple_of=8), n_experts=4)
    mx.reset_parameters()
    lg, emb = mx(x, include_embeds=True)
    assert lg.shape == (2, 9, 50) and torch.isfinite(lg).all()


def test_get_model_without_checkpoint_initialises_every_family():
    """No HF checkpoint -> the registered variant is materialised from the meta device and must be INITIALISED (not the
    uninitialised storage ``to_empty`` leaves behind) for all three families."""
    from speculator.train_speculator_utils import get_model, register_model
    register_model("embedgpt_bigcode", "test", lambda: EmbedGPTBigCode(vocab=50, emb_dim=32, nheads=4, nlayers=2, max_pos=64))
    register_model("embedmixtral", "test", lambda: EmbedMixtral(
        LLaMAConfig(src_vocab_size=50, emb_dim=32, nheads=4, kvheads=2, nlayers=2, multiple_of=8), n_experts=4))
    x = torch.randint(0, 50, (2, 9))
    for arch, variant in (("embedgpt_bigcode", "test"), ("embedmixtral", "test"), ("embedllama", "tiny")):
        m = get_model(arch, variant, model_path=None, device_type="cpu", dtype=torch.float32)
        assert all(torch.isfinite(p).all() and p.abs().max() < 10 for p in m.parameters()), arch
        assert not any(p.requires_grad for p in m.parameters())
        lg = m(x)
        assert torch.isfinite(lg).all() and lg.std() > 0, arch


def test_hf_loader_roundtrip():
    import fms_to_hf_llama as ex
    from fms_fsdp_b200.models.hf_loader import load_hf_llama
    from fms_fsdp_b200.models.llama import LLaMA
    torch.manual_seed(3)
    m = LLaMA(LLaMAConfig(src_vocab_size=40, emb_dim=32, nheads=4, kvheads=2, nlayers=2, multiple_of=8, max_expected_seq_len=64))
    m.reset_parameters()
    d = tempfile.mkdtemp()
    ex.convert_to_hf(m, "llama2_x").save_pretrained(d)
    back = load_hf_llama(d, dtype=torch.float32)
    x = torch.randint(0, 40, (2, 11))
    assert torch.allclose(m(x), back(x), atol=1e-5)


def test_hf_gpt_bigcode_and_mixtral_checkpoints_load_with_equal_logits():
    """``get_model(..., model_path=<HF dir>)`` loads GPT-BigCode and Mixtral weights (reference adapters
    ``train_speculator_utils.py:526-569``): logits equal the ``transformers`` implementation on the same checkpoint."""
    from transformers import GPTBigCodeConfig, GPTBigCodeForCausalLM, MixtralConfig, MixtralForCausalLM
    from speculator.train_speculator_utils import get_model
    torch.manual_seed(5)
    x = torch.randint(0, 50, (2, 13))
    d1 = tempfile.mkdtemp()
    hf = GPTBigCodeForCausalLM(GPTBigCodeConfig(vocab_size=50, n_positions=64, n_embd=32, n_layer=2, n_head=4, n_inner=128,
                                                multi_query=True, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)).eval()
    hf.save_pretrained(d1)
    ours = get_model("embedgpt_bigcode", "20b", model_path=d1, device_type="cpu", dtype=torch.float32)
    with torch.no_grad():
        assert torch.allclose(ours(x), hf(x).logits, atol=2e-4)
    d2 = tempfile.mkdtemp()
    hf2 = MixtralForCausalLM(MixtralConfig(vocab_size=50, hidden_size=32, intermediate_size=48, num_hidden_layers=2,
                                           num_attention_heads=4, num_key_value_heads=2, num_local_experts=4,
                                           num_experts_per_tok=2, max_position_embeddings=64)).eval()
    hf2.save_pretrained(d2)
    ours2 = get_model("embedmixtral", "8x7b", model_path=d2, device_type="cpu", dtype=torch.float32)
    with torch.no_grad():
        assert torch.allclose(ours2(x), hf2(x).logits, atol=2e-4)


def test_speculator_entrypoint_resumes_from_its_checkpoint(tmp_path):
    """Stop after 4 steps, start again with ``num_steps=6`` on the same checkpoint folder: the run picks up speculator,
    optimizer and step count (reference ``train_speculator.py:246-262``) and only trains steps 5 and 6."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [sys.executable, os.path.join(root, "speculator", "train_speculator.py"), "--model_arch=embedllama",
            "--model_variant=tiny", "--model_path=/nonexistent", "--sharding_strategy=ddp", "--use_dummy_dataset=True",
            "--report_interval=1", "--stage2_start_step=100", "--n_speculator_heads=2", "--speculator_width=32",
            "--seq_length=16", "--vocab_size=512", "--batch_size=2", f"--ckpt_save_path={tmp_path}",
            f"--ckpt_load_path={tmp_path}", "--checkpoint_interval=100", "--comm_backend=gloo", "--use_torch_compile=False"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["OMP_NUM_THREADS"] = "1"
    r1 = subprocess.run(base + ["--num_steps=4"], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r1.returncode == 0, r1.stdout[-1500:] + r1.stderr[-1500:]
    assert "step: 4" in r1.stdout and os.path.isdir(os.path.join(tmp_path, "checkpoints", "step_4_ckp"))
    r2 = subprocess.run(base + ["--num_steps=6"], capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-1500:]
    assert "Prior checkpoint" in r2.stdout and "step_4_ckp" in r2.stdout
    assert "step: 5" in r2.stdout and "step: 6" in r2.stdout and "step: 3" not in r2.stdout
    assert os.path.isdir(os.path.join(tmp_path, "checkpoints", "step_6_ckp"))


def test_hf_loader_refuses_what_it_cannot_reproduce():
    """Rope scaling types that are not implemented (yarn, dynamic), sliding-window attention and projection biases would load
    without error and silently change the logits: the loader must stop instead.  ``llama3`` / ``linear`` scaling is carried
    into the model config (``tests/test_exporters.py`` checks the logits)."""
    from fms_fsdp_b200.models.hf_loader import config_from_hf
    base = dict(vocab_size=64, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                max_position_embeddings=128, rope_theta=10000.0)
    assert config_from_hf(base).emb_dim == 32
    assert config_from_hf({**base, "rope_scaling": None, "sliding_window": None}).nlayers == 1
    v5 = {k: v for k, v in base.items() if k != "rope_theta"}       # transformers 5 keeps theta inside rope_parameters
    assert config_from_hf({**v5, "rope_parameters": {"rope_type": "default", "rope_theta": 5e5}}).rope_theta == 5e5
    assert config_from_hf({**base, "rope_scaling": {"rope_type": "llama3", "factor": 8.0}}).rope_scaling["factor"] == 8.0
    assert config_from_hf({**base, "rope_scaling": {"type": "linear", "factor": 2.0}}).rope_scaling["type"] == "linear"
    for bad, msg in (({"rope_scaling": {"rope_type": "dynamic", "factor": 2.0}}, "rope scaling"),
                     ({"rope_parameters": {"rope_type": "yarn", "rope_theta": 1e4}}, "rope scaling"),
                     ({"sliding_window": 64}, "sliding-window"), ({"attention_bias": True}, "biases")):
        with pytest.raises(NotImplementedError, match=msg):
            config_from_hf({**base, **bad})


def test_greedy_generation_equals_transformers_generate(tmp_path):
    """Prefill + KV-cache decode of the frozen base model against an independent decoder: the same HF Llama weights, greedy
    decoding, ``transformers.generate`` vs ``generate(..., use_cache=True)`` and ``use_cache=False`` -- identical tokens."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from speculator.train_speculator_utils import get_model
    torch.manual_seed(11)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=128, hidden_size=64, intermediate_size=96, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=128,
                                      rope_theta=10000.0, tie_word_embeddings=False)).eval()
    d = str(tmp_path / "hf")
    hf.save_pretrained(d)
    ours = get_model("embedllama", "7b", model_path=d, device_type="cpu", dtype=torch.float32)
    prompt = torch.randint(0, 128, (3, 9))
    with torch.no_grad():
        want = hf.generate(prompt, max_new_tokens=14, do_sample=False, pad_token_id=0)
        got, embeds = generate(ours, prompt, max_new_tokens=14, do_sample=False, use_cache=True)
        got2 = generate(ours, prompt, max_new_tokens=14, do_sample=False, use_cache=False, include_embeds=False)
    assert torch.equal(got, want) and torch.equal(got2, want)
    assert embeds.shape[0] == 3 and embeds.shape[-1] == 64


def test_mixtral_loader_reads_the_fused_in_memory_expert_layout(tmp_path):
    """Besides the hub layout (``block_sparse_moe.experts.N.w1/w2/w3``, what ``save_pretrained`` writes) the loader accepts the
    fused tensors of recent ``transformers`` state dicts (``mlp.experts.gate_up_proj`` / ``down_proj``) -- same logits."""
    import json
    from safetensors.torch import save_file
    from transformers import MixtralConfig, MixtralForCausalLM
    from speculator.train_speculator_utils import get_model
    torch.manual_seed(8)
    cfg = MixtralConfig(vocab_size=50, hidden_size=32, intermediate_size=48, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, num_local_experts=4, num_experts_per_tok=2, max_position_embeddings=64)
    hf = MixtralForCausalLM(cfg).eval()
    sd = {k: v.detach().clone().contiguous() for k, v in hf.state_dict().items()}
    if not any(k.endswith("mlp.experts.gate_up_proj") for k in sd):
        pytest.skip("this transformers version keeps per-expert tensors in memory")
    os.makedirs(tmp_path / "m")
    save_file(sd, str(tmp_path / "m" / "model.safetensors"))
    with open(tmp_path / "m" / "config.json", "w") as f:
        json.dump(cfg.to_dict(), f)
    ours = get_model("embedmixtral", "8x7b", model_path=str(tmp_path / "m"), device_type="cpu", dtype=torch.float32)
    x = torch.randint(0, 50, (2, 13))
    with torch.no_grad():
        assert torch.allclose(ours(x), hf(x).logits, atol=2e-4)


def test_hf_loader_reads_pytorch_bin_checkpoints(tmp_path):
    """Older HF checkpoints ship ``pytorch_model.bin`` instead of safetensors: same loader, same logits."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from fms_fsdp_b200.models.hf_loader import load_hf_llama
    torch.manual_seed(9)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=60, hidden_size=32, intermediate_size=48, num_hidden_layers=1,
                                      num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=64,
                                      tie_word_embeddings=True)).eval()          # tied head: no lm_head tensor on disk
    d = str(tmp_path / "m")
    os.makedirs(d)
    torch.save({k: v for k, v in hf.state_dict().items() if k != "lm_head.weight"}, os.path.join(d, "pytorch_model.bin"))
    hf.config.save_pretrained(d)
    ours = load_hf_llama(d, "cpu", torch.float32).eval()
    x = torch.randint(0, 60, (2, 12))
    with torch.no_grad():
        assert torch.allclose(ours(x), hf(x).logits, atol=2e-4)
    empty = str(tmp_path / "empty")
    os.makedirs(empty)
    hf.config.save_pretrained(empty)
    with pytest.raises(FileNotFoundError, match="no safetensors"):
        load_hf_llama(empty, "cpu", torch.float32)
