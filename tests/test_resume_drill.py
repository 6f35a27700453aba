"""End-to-end recovery drill of the Llama entry point on CPU (gloo, 2 ranks): a rank is killed mid-run with
``--fault_inject_step``, the job is simply restarted, auto-resume picks up the newest checkpoint of the
save directory and the run finishes -- the reference's recovery model (restart + auto-resume), exercised for real."""
import os
import re
import subprocess
import sys

import pytest
from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(tmp_path, port, *extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "main_training_llama.py"), "--model_variant=llama2_tiny",
           "--use_dummy_dataset=True", "--sharding_strategy=fsdp", "--report_interval=1", "--seq_length=32",
           "--vocab_size=512", "--batch_size=2", f"--ckpt_save_path={tmp_path}", f"--ckpt_load_path={tmp_path}",
           "--checkpoint_interval=2", "--comm_backend=gloo", "--use_torch_compile=False", *extra]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def _steps(stdout):
    return [int(m) for m in re.findall(r"^step: (\d+)$", stdout, flags=re.M)]


@pytest.mark.timeout(1500)
def test_kill_a_rank_restart_and_auto_resume(tmp_path):
    # 1) rank 1 dies at step 5; checkpoints of steps 2 and 4 are already on disk
    r1 = _launch(tmp_path, free_port(), "--num_steps=8", "--fault_inject_step=5")
    assert r1.returncode != 0, "the injected fault must take the job down"
    assert "[fault-inject]" in (r1.stdout + r1.stderr)
    ck = sorted(os.listdir(os.path.join(tmp_path, "checkpoints")))
    assert any(d.startswith("step_4") for d in ck), ck
    # 2) plain restart: resumes from step 4 (not from scratch) and completes
    r2 = _launch(tmp_path, free_port(), "--num_steps=8")
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    steps = _steps(r2.stdout)
    assert steps and steps[0] == 5 and steps[-1] == 8, steps
    assert "Checkpoint saved" in r2.stdout


@pytest.mark.timeout(1500)
@pytest.mark.parametrize("strategy,nproc,extra", [("hsdp", 4, ["--hsdp_shard_size=2"]), ("ddp", 2, [])])
def test_entrypoint_hsdp_and_ddp_train_checkpoint_and_resume(tmp_path, strategy, nproc, extra):
    """The other two data-parallel layouts through the public entry point (gloo): HSDP as a 2x2 mesh, DDP; each trains,
    writes its checkpoint with the right writer election, and a restart resumes from it."""
    def launch(num_steps):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
               "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "main_training_llama.py"),
               "--model_variant=llama2_tiny", "--use_dummy_dataset=True", f"--sharding_strategy={strategy}", *extra,
               "--report_interval=1", "--seq_length=32", "--vocab_size=512", "--batch_size=2", f"--ckpt_save_path={tmp_path}",
               f"--ckpt_load_path={tmp_path}", "--checkpoint_interval=100", "--comm_backend=gloo", f"--num_steps={num_steps}"]
        return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    r1 = launch(2)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    assert _steps(r1.stdout) == [1, 2]
    want_mesh = "mesh=replica2xshard2" if strategy == "hsdp" else f"mesh=replica{nproc}xshard1"
    assert want_mesh in r1.stdout, r1.stdout[-1500:]
    r2 = launch(4)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert _steps(r2.stdout) == [3, 4] and "Prior checkpoint" in r2.stdout


@pytest.mark.timeout(900)
def test_entrypoint_policy_flags_low_cpu_init_selective_recompute_fp32(tmp_path):
    """``--low_cpu_fsdp`` (meta-device construction, unit-at-a-time materialisation, rank 0's init broadcast),
    ``--fsdp_activation_checkpointing --selective_checkpointing=1/2`` and ``--mixed_precision=False`` together, 2 ranks."""
    r = _launch(tmp_path, free_port(), "--num_steps=2", "--low_cpu_fsdp=True", "--fsdp_activation_checkpointing=True",
                "--selective_checkpointing=1/2", "--mixed_precision=False", "--checkpoint_interval=100")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert _steps(r.stdout) == [1, 2] and "applying FSDP activation checkpointing" in r.stdout
    losses = [float(x) for x in re.findall(r"^loss: ([0-9.]+)$", r.stdout, flags=re.M)]
    assert len(losses) == 2 and all(5.0 < v < 9.0 for v in losses), losses      # ~ln(vocab) at initialisation, finite
