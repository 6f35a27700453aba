"""The Llama entry point on a real (tiny) arrow corpus through the full stateful loader stack -- sampling over two
datasets, packing to seq_length, worker auto-checkpointing -- then a restart that resumes BOTH model and loader state
(reference call stack: SURVEY.md 3.1 / 3.3 / 3.4)."""
import os
import re
import subprocess
import sys

import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(path, docs):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    schema = pa.schema([pa.field("tokens", pa.uint32())])
    with pa.ipc.new_file(path, schema) as w:
        for d in docs:
            w.write(pa.record_batch([pa.array(d, type=pa.uint32())], schema=schema))


def _corpus(tmp):
    _write(f"{tmp}/dataset_1/a.arrow", [[(7 * i + j) % 500 + 1 for j in range(90)] for i in range(60)])
    _write(f"{tmp}/dataset_2/b.arrow", [[(11 * i + 3 * j) % 500 + 1 for j in range(40)] for i in range(80)])
    os.makedirs(f"{tmp}/meta")
    with open(f"{tmp}/meta/combined_counts.csv", "w") as f:
        f.write("dataset/filename,documents,tokens\n/dataset_1/a.arrow,60,5400\n/dataset_2/b.arrow,80,3200\n")


def _run(data, ckpt, steps):
    cmd = [sys.executable, os.path.join(ROOT, "main_training_llama.py"), "--model_variant=llama2_tiny",
           "--use_dummy_dataset=False", f"--data_path={data}", "--datasets=dataset_1,dataset_2", "--weights=2,1",
           "--file_type=arrow", "--col_name=tokens", "--logical_shards=8", "--num_workers=1", "--seq_length=32",
           "--vocab_size=512", "--batch_size=2", "--eos_token=0", f"--num_steps={steps}", "--report_interval=1",
           "--checkpoint_interval=3", f"--ckpt_save_path={ckpt}", f"--ckpt_load_path={ckpt}", "--sharding_strategy=fsdp",
           "--comm_backend=gloo", "--use_torch_compile=False"]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))


def test_arrow_corpus_train_checkpoint_and_resume(tmp_path):
    data, ckpt = str(tmp_path / "data"), str(tmp_path / "ckpt")
    _corpus(data)
    r1 = _run(data, ckpt, 3)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    assert "step: 3" in r1.stdout and "Checkpoint saved" in r1.stdout
    step_dir = os.path.join(ckpt, "checkpoints", "step_3_ckp")
    assert os.path.isdir(step_dir) and any(f.startswith("loader_state") for f in os.listdir(step_dir)), os.listdir(step_dir)
    r2 = _run(data, ckpt, 5)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    steps = [int(s) for s in re.findall(r"^step: (\d+)$", r2.stdout, flags=re.M)]
    assert steps == [4, 5], steps
    losses = [float(x) for x in re.findall(r"^loss: ([0-9.eE+-]+)$", r1.stdout + r2.stdout, flags=re.M)]
    assert len(losses) == 5 and all(l == l and l < 20 for l in losses), losses


def test_new_run_from_another_runs_weights_with_and_without_its_data_position(tmp_path):
    """``--ckpt_load_path=<run A> --ckpt_save_path=<run B>``: B starts at step 0 from A's weights.  With
    ``--resuming_dataset=True`` the loader additionally continues from A's position; by default the data starts over."""
    data, a = str(tmp_path / "data"), str(tmp_path / "A")
    _corpus(data)
    assert _run(data, a, 3).returncode == 0

    def start_b(name, *extra):
        cmd = [sys.executable, os.path.join(ROOT, "main_training_llama.py"), "--model_variant=llama2_tiny",
               "--use_dummy_dataset=False", f"--data_path={data}", "--datasets=dataset_1,dataset_2", "--weights=2,1",
               "--file_type=arrow", "--col_name=tokens", "--logical_shards=8", "--num_workers=1", "--seq_length=32",
               "--vocab_size=512", "--batch_size=2", "--eos_token=0", "--num_steps=2", "--report_interval=1",
               "--checkpoint_interval=100", f"--ckpt_save_path={tmp_path / name}", f"--ckpt_load_path={a}",
               "--sharding_strategy=fsdp", "--comm_backend=gloo", *extra]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert [int(s) for s in re.findall(r"^step: (\d+)$", r.stdout, flags=re.M)] == [1, 2]
        assert "Prior checkpoint" in r.stdout and "step_3_ckp" in r.stdout
        return r.stdout

    with_data = start_b("B1", "--resuming_dataset=True")
    assert "Dataset checkpoint loaded" in with_data
    fresh_data = start_b("B2")
    assert "Dataset checkpoint loaded" not in fresh_data and "dataset starting from scratch" in fresh_data
