"""docs/fine_tuning.md end to end on CPU: public HF Llama weights -> ``hf_to_fms_llama.py`` -> ``main_training_llama.py
--ckpt_load_path=<file>`` (fresh optimizer, step 0) -> annealing phase from the new run's checkpoint -> HF export whose logits
differ from the starting point only by the few training steps."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, **kw):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["OMP_NUM_THREADS"] = "1"
    r = subprocess.run([sys.executable] + cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env, **kw)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_hf_weights_to_continued_training_to_annealing_to_export(tmp_path):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    hf = LlamaForCausalLM(LlamaConfig(vocab_size=1024, hidden_size=256, intermediate_size=704, num_hidden_layers=2,
                                      num_attention_heads=4, num_key_value_heads=4, max_position_embeddings=512,
                                      rope_theta=10000.0, rms_norm_eps=1e-5, tie_word_embeddings=False)).eval()
    hf_dir, pth, run1 = str(tmp_path / "hf"), str(tmp_path / "start.pth"), str(tmp_path / "run1")
    hf.save_pretrained(hf_dir)
    out = _run(["hf_to_fms_llama.py", f"--hf_path={hf_dir}", f"--save_path={pth}", "--model_variant=llama2_tiny", "--dtype=fp32"])
    assert "wrote" in out and os.path.isfile(pth)

    common = ["main_training_llama.py", "--model_variant=llama2_tiny", "--use_dummy_dataset=True", "--sharding_strategy=fsdp",
              "--report_interval=1", "--seq_length=32", "--vocab_size=1024", "--batch_size=2", "--checkpoint_interval=100",
              "--comm_backend=gloo", "--learning_rate=1e-4"]
    out = _run(common + ["--num_steps=2", f"--ckpt_load_path={pth}", f"--ckpt_save_path={run1}"])
    assert "single-file checkpoint" in out and "step: 1" in out and "step: 2" in out
    assert os.path.isdir(os.path.join(run1, "checkpoints", "step_2_ckp"))

    # annealing phase: a NEW run directory that starts from run1's checkpoint (not a resume: step counter restarts)
    run2 = str(tmp_path / "run2")
    out = _run(common + ["--num_steps=3", "--training_stage=annealing", f"--ckpt_load_path={run1}/checkpoints/step_2_ckp",
                         f"--ckpt_save_path={run2}"])
    assert "Prior checkpoint" in out and "step: 1" in out and "step: 3" in out
    assert os.path.isdir(os.path.join(run2, "checkpoints", "step_3_ckp"))

    exp = str(tmp_path / "export")
    _run(["fms_to_hf_llama.py", "--model_variant=llama2_tiny", f"--load_path={run2}/checkpoints/step_3_ckp", f"--save_path={exp}"])
    tuned = LlamaForCausalLM.from_pretrained(exp).eval()
    x = torch.randint(0, 1024, (2, 16))
    with torch.no_grad():
        a, b = hf(x).logits, tuned(x).logits
    drift = (a - b).abs().max().item()
    assert 0 < drift < 0.5 * a.abs().max().item(), drift       # trained a little, still the same model
