"""Mamba2-hybrid model: engine (tuple residual stream across unit boundaries, selective recompute) vs plain autograd;
chunked SSD form vs sequential oracle; entry point smoke on CPU."""
import copy
import subprocess
import sys
import os

import pytest
import torch

from fms_fsdp_b200.models.mamba import Block, MambaConfig, MambaLMHeadModel
from fms_fsdp_b200.ops import torch_kernels as TK
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel
from fms_fsdp_b200.policies import apply_fsdp_checkpointing
from fms_fsdp_b200.utils.config_utils import get_model_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny():
    torch.manual_seed(0)
    m = MambaLMHeadModel(MambaConfig(**get_model_config("mamba_tiny")))
    m.reset_parameters()
    return m


def test_structure_and_state_dict_names():
    m = _tiny()
    kinds = [type(b.mixer).__name__ for b in m.backbone.layers]
    assert kinds == ["Mamba2", "Mamba2", "MHA", "Mamba2"]
    keys = set(m.state_dict())
    for k in ["backbone.embedding.weight", "backbone.layers.0.mixer.in_proj.weight", "backbone.layers.0.mixer.conv1d.weight",
              "backbone.layers.0.mixer.A_log", "backbone.layers.0.mixer.dt_bias", "backbone.layers.0.mixer.D",
              "backbone.layers.0.mixer.norm.weight", "backbone.layers.0.mlp.fc1.weight", "backbone.layers.2.mixer.out_proj.weight",
              "backbone.norm_f.weight", "lm_head.weight"]:
        assert k in keys, k
    assert m.state_dict()["backbone.layers.0.mixer.conv1d.weight"].shape[1:] == (1, 4)
    assert m.lm_head.weight.shape[0] % 16 == 0
    out = m(torch.randint(0, 512, (1, 16)))
    assert hasattr(out, "logits") and out.logits.shape == (1, 16, 512)


@pytest.mark.parametrize("ac", [None, "1/2"])
def test_engine_matches_plain_autograd(ac):
    m = _tiny()
    ref = copy.deepcopy(m)
    if ac:
        apply_fsdp_checkpointing(m, Block, ac)
    eng = ShardedModel(m, device="cpu"); opt = ShardedAdamW(eng, lr=2e-3)
    ropt = torch.optim.AdamW(ref.parameters(), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.1)
    x = torch.randint(0, 512, (2, 48))
    for _ in range(3):
        l = eng.forward_backward(x, x); gn = eng.clip_grad_norm_(1.0); opt.step()
        ropt.zero_grad(); rl = ref(x, labels=x); rl.backward()
        rgn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0); ropt.step()
        assert l.item() == pytest.approx(rl.item(), rel=1e-5) and gn.item() == pytest.approx(rgn.item(), rel=1e-4)


def test_ssd_chunked_equals_sequential_with_padding():
    torch.manual_seed(1)
    S, H, P, G, N = 70, 4, 8, 2, 16
    x = torch.randn(2 * S, H, P); dt = torch.randn(2 * S, H); A = -torch.rand(H) - 0.1
    Bm = torch.randn(2 * S, G, N); Cm = torch.randn(2 * S, G, N); D = torch.randn(H); b = torch.randn(H)
    assert torch.allclose(TK.ssd_scan_fwd(x, dt, A, Bm, Cm, D, b, S, 16), TK.ssd_scan_chunked(x, dt, A, Bm, Cm, D, b, S, 16), atol=1e-4)


def test_selective_scan_op_grads():
    from fms_fsdp_b200 import ops
    torch.manual_seed(2)
    S, Dm, N = 12, 6, 4
    u = torch.randn(S, Dm, requires_grad=True); delta = torch.randn(S, Dm, requires_grad=True)
    A = torch.nn.Parameter(-torch.rand(Dm, N)); Bm = torch.randn(S, N, requires_grad=True); Cm = torch.randn(S, N, requires_grad=True)
    Dp = torch.nn.Parameter(torch.randn(Dm)); z = torch.randn(S, Dm, requires_grad=True)
    y = ops.selective_scan(u, delta, A, Bm, Cm, Dp, z, None, S)
    y.square().sum().backward()
    assert all(t.grad is not None and torch.isfinite(t.grad).all() for t in (u, delta, A, Bm, Cm, Dp, z))


def test_mamba_entrypoint_resume_and_export(tmp_path):
    """Train 3 steps, restart to 5 (auto-resume: only steps 4 and 5 run), then export the final checkpoint to the mamba_ssm /
    HF directory format with ``fms_to_hf_mamba.py`` -- the whole life cycle through the public CLIs."""
    base = [sys.executable, os.path.join(ROOT, "main_training_mamba.py"), "--model_variant=mamba_tiny",
            "--use_dummy_dataset=True", "--sharding_strategy=fsdp", "--report_interval=1", "--seq_length=32",
            "--vocab_size=512", f"--ckpt_save_path={tmp_path}", f"--ckpt_load_path={tmp_path}", "--checkpoint_interval=100",
            "--comm_backend=gloo"]
    r1 = subprocess.run(base + ["--num_steps=3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r1.returncode == 0, r1.stdout[-1500:] + r1.stderr[-1500:]
    assert "step: 3" in r1.stdout and "Checkpoint saved" in r1.stdout
    r2 = subprocess.run(base + ["--num_steps=5"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-1500:]
    assert "Prior checkpoint" in r2.stdout and "step: 4" in r2.stdout and "step: 5" in r2.stdout and "step: 2" not in r2.stdout
    out = os.path.join(tmp_path, "hf")
    r3 = subprocess.run([sys.executable, os.path.join(ROOT, "fms_to_hf_mamba.py"), "--model_variant=mamba_tiny",
                         f"--load_path={tmp_path}/checkpoints/step_5_ckp", f"--save_path={out}"],
                        capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r3.returncode == 0, r3.stdout[-1500:] + r3.stderr[-1500:]
    assert os.path.exists(os.path.join(out, "config.json")) and any(f.endswith((".bin", ".safetensors")) for f in os.listdir(out))


def test_mamba2_matches_the_transformers_implementation():
    """An independent implementation of the same architecture: ``transformers.Mamba2ForCausalLM`` (its pure-torch path) loaded
    with OUR state dict (one key renamed) produces the same logits and the same gradients -- conv1d, SSD scan incl. a sequence
    that is not a multiple of the chunk size, D skip, dt bias + softplus, gated RMSNorm, fp32 residual stream."""
    from transformers import Mamba2Config, Mamba2ForCausalLM
    from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
    torch.manual_seed(0)
    cfg = MambaConfig(d_model=128, d_intermediate=0, n_layer=2, vocab_size=512,
                      ssm_cfg={"layer": "Mamba2", "headdim": 32, "d_state": 32, "chunk_size": 32}, attn_layer_idx=[], attn_cfg={},
                      rms_norm=True, residual_in_fp32=True, fused_add_norm=True, pad_vocab_size_multiple=16, tie_embeddings=False)
    ours = MambaLMHeadModel(cfg); ours.reset_parameters()
    with torch.no_grad():
        for n, p in ours.named_parameters():
            if n.endswith(("conv1d.bias", ".D", "dt_bias")):
                p.normal_(0, 0.3)
    hf = Mamba2ForCausalLM(Mamba2Config(
        vocab_size=512, hidden_size=128, state_size=32, num_hidden_layers=2, head_dim=32, num_heads=8, expand=2, n_groups=1,
        conv_kernel=4, chunk_size=32, tie_word_embeddings=False, rms_norm=True, use_bias=False, use_conv_bias=True,
        residual_in_fp32=True, layer_norm_epsilon=1e-5, hidden_act="silu"))
    hf.load_state_dict({("backbone.embeddings.weight" if k == "backbone.embedding.weight" else k): v.clone()
                        for k, v in ours.state_dict().items()}, strict=True)
    x = torch.randint(0, 512, (2, 48))
    a = ours(x)
    a = a.logits if hasattr(a, "logits") else a
    b = hf(x).logits
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-4), (a - b).abs().max()
    w = torch.randn_like(a)
    (a * w).sum().backward()
    (b * w).sum().backward()
    theirs = dict(hf.named_parameters())
    for n, p in ours.named_parameters():
        q = theirs["backbone.embeddings.weight" if n == "backbone.embedding.weight" else n]
        err, scale = (p.grad - q.grad).abs().max().item(), q.grad.abs().max().item()
        assert err <= 1e-4 * scale + 1e-6, (n, err, scale)          # fp32 summation-order noise only


def test_hybrid_matches_transformers_bamba_and_partial_rotary_matches_the_formula():
    """The Mamba2 + attention (GQA, rotary) + gated-MLP hybrid against ``transformers.BambaForCausalLM`` carrying our weights
    (``fms_to_hf_mamba.to_transformers``): same logits, same gradients.  transformers 5.5 builds full-width rotary tables for
    Bamba, so that comparison uses ``rotary_emb_dim = head_dim``; the partial rotary of ``mamba_9.8b`` (first half of each head,
    GPT-NeoX pairing) is checked against the written-out formula."""
    import fms_to_hf_mamba as ex
    from fms_fsdp_b200.models.mamba import MambaConfig, MambaLMHeadModel
    from fms_fsdp_b200.utils.config_utils import get_model_config
    d = get_model_config("mamba_tiny")
    d["attn_cfg"]["rotary_emb_dim"] = d["attn_cfg"]["head_dim"]
    torch.manual_seed(0)
    ours = MambaLMHeadModel(MambaConfig(**d)); ours.reset_parameters()
    with torch.no_grad():
        for n, p in ours.named_parameters():
            if n.endswith(("conv1d.bias", ".D", "dt_bias")):
                p.normal_(0, 0.3)
    hf = ex.to_transformers(ours)
    assert type(hf).__name__ == "BambaForCausalLM"
    x = torch.randint(0, 512, (2, 48))
    a = ours(x)
    a = a.logits if hasattr(a, "logits") else a
    b = hf(x).logits
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-4), (a - b).abs().max()
    w = torch.randn_like(a)
    (a * w).sum().backward()
    (b * w).sum().backward()
    g = {n: p.grad for n, p in hf.named_parameters()}
    blk = ours.backbone.layers[2]                                   # the attention block
    H, KV, hd = 4, 2, 32
    q, k, v = blk.mixer.in_proj.weight.grad.split([H * hd, KV * hd, KV * hd])
    for mine, theirs in ((q, "self_attn.q_proj"), (k, "self_attn.k_proj"), (v, "self_attn.v_proj"),
                         (blk.mixer.out_proj.weight.grad, "self_attn.o_proj"), (blk.mlp.fc2.weight.grad, "feed_forward.down_proj"),
                         (blk.mlp.fc1.weight.grad.chunk(2)[1], "feed_forward.gate_proj")):
        t = g[f"model.layers.2.{theirs}.weight"]
        assert (mine - t).abs().max() <= 1e-4 * t.abs().max() + 1e-6, theirs
    t = g["model.layers.0.mamba.in_proj.weight"]
    assert (ours.backbone.layers[0].mixer.in_proj.weight.grad - t).abs().max() <= 1e-4 * t.abs().max() + 1e-6

    # partial rotary (rotary_emb_dim = head_dim / 2, the mamba_9.8b setting) vs the formula
    torch.manual_seed(1)
    att = MambaLMHeadModel(MambaConfig(**get_model_config("mamba_tiny"))).backbone.layers[2].mixer
    att.reset_parameters()
    h = torch.randn(2, 40, 128)
    rd = 16
    with torch.no_grad():
        y = att(h)
        qq, kk, vv = att.in_proj(h).split([H * hd, KV * hd, KV * hd], -1)
        qq, kk, vv = (t.view(2, 40, -1, hd).transpose(1, 2) for t in (qq, kk, vv))
        ang = torch.outer(torch.arange(40).float(), 1.0 / (10000.0 ** (torch.arange(0, rd, 2).float() / rd)))
        cos, sin = torch.cat([ang.cos(), ang.cos()], -1), torch.cat([ang.sin(), ang.sin()], -1)

        def rot(t):
            r, keep = t[..., :rd], t[..., rd:]
            return torch.cat([r * cos + torch.cat([-r[..., rd // 2:], r[..., :rd // 2]], -1) * sin, keep], -1)
        o = torch.nn.functional.scaled_dot_product_attention(rot(qq), rot(kk), vv, is_causal=True, enable_gqa=True)
        want = att.out_proj(o.transpose(1, 2).reshape(2, 40, H * hd))
    assert torch.allclose(y, want, atol=1e-5, rtol=1e-4), (y - want).abs().max()


def test_mamba1_layer_type_matches_transformers_and_trains_through_the_engine():
    """``ssm_cfg = {"layer": "Mamba1"}`` (mamba_ssm's default layer type): the selective-scan mixer equals
    ``transformers.MambaForCausalLM`` with the same weights (logits + every gradient), the sharded runtime reproduces plain
    autograd + AdamW on it, and the exporter writes the ``transformers`` layout."""
    import fms_to_hf_mamba as ex
    torch.manual_seed(0)
    ours = MambaLMHeadModel(MambaConfig(**get_model_config("mamba1_tiny"))); ours.reset_parameters()
    assert [type(b.mixer).__name__ for b in ours.backbone.layers] == ["Mamba1"] * 3
    hf = ex.to_transformers(ours)
    assert type(hf).__name__ == "MambaForCausalLM"
    x = torch.randint(0, 512, (2, 40))
    a, b = ours(x).logits, hf(x).logits
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-4), (a - b).abs().max()
    w = torch.randn_like(a)
    (a * w).sum().backward()
    (b * w).sum().backward()
    theirs = dict(hf.named_parameters())
    for n, p in ours.named_parameters():
        q = theirs["backbone.embeddings.weight" if n == "backbone.embedding.weight" else n]
        assert (p.grad - q.grad).abs().max() <= 1e-4 * q.grad.abs().max() + 1e-6, n
    ours.zero_grad()

    ref = copy.deepcopy(ours)
    eng = ShardedModel(ours, device="cpu"); opt = ShardedAdamW(eng, lr=2e-3)
    ropt = torch.optim.AdamW(ref.parameters(), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.1)
    for _ in range(3):
        l = eng.forward_backward(x, x); gn = eng.clip_grad_norm_(1.0); opt.step()
        ropt.zero_grad(); rl = ref(x, labels=x); rl.backward()
        rgn = torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0); ropt.step()
        assert l.item() == pytest.approx(rl.item(), rel=1e-5) and gn.item() == pytest.approx(rgn.item(), rel=1e-4)
    with pytest.raises(ValueError, match="Invalid ssm_layer"):
        MambaLMHeadModel(MambaConfig(**{**get_model_config("mamba1_tiny"), "ssm_cfg": {"layer": "Mamba3"}}))
