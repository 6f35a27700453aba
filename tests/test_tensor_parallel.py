"""Tensor-parallel frozen base model of speculator training (reference: fms TP strategy, `train_speculator.py:133-160`):
slicing a loaded LLaMA over 2 ranks (gloo) must reproduce the unsharded logits, embeds and KV-cache decode."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import free_port
from fms_fsdp_b200.parallel.tensor_parallel import shard_llama_for_tp
from fms_fsdp_b200.utils.config_utils import get_model_config


def _worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speculator"))
        from train_speculator_utils import EmbedLLaMA
        torch.manual_seed(0)
        m = EmbedLLaMA(get_model_config("llama2_tiny")); m.reset_parameters(); m.eval()
        x = torch.randint(0, m.config.src_vocab_size, (2, 12))
        with torch.no_grad():
            ref_logits, ref_embeds = m(x, include_embeds=True)
            # prefill + one cached decode step, unsharded
            l0, cache = m(x[:, :-1], use_cache=True)
            l1, _ = m(x[:, -1:], past_key_value_states=cache, use_cache=True)
            shard_llama_for_tp(m, dist.group.WORLD)
            tp_logits, tp_embeds = m(x, include_embeds=True)
            t0, tcache = m(x[:, :-1], use_cache=True)
            t1, _ = m(x[:, -1:], past_key_value_states=tcache, use_cache=True)
        ok = dict(logits=(tp_logits - ref_logits).abs().max().item(), embeds=(tp_embeds - ref_embeds).abs().max().item(),
                  prefill=(t0 - l0).abs().max().item(), decode=(t1 - l1).abs().max().item(),
                  scale=ref_logits.abs().max().item())
        if rank == 0:
            torch.save(ok, os.path.join(outdir, "out.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_tp2_base_model_matches_unsharded():
    outdir = tempfile.mkdtemp()
    mp.spawn(_worker, args=(2, free_port(), outdir), nprocs=2, join=True)
    r = torch.load(os.path.join(outdir, "out.pt"), weights_only=False)
    tol = 1e-4 * max(1.0, r["scale"])
    assert r["logits"] < tol and r["embeds"] < tol and r["prefill"] < tol and r["decode"] < tol, r


def _family_worker(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "speculator"))
        from train_speculator_utils import EmbedGPTBigCode, EmbedMixtral
        from fms_fsdp_b200.models.llama import LLaMAConfig
        from fms_fsdp_b200.parallel.tensor_parallel import shard_for_tp
        res = {}
        for name in ("gpt_bigcode", "mixtral"):
            torch.manual_seed(1)
            if name == "gpt_bigcode":
                m = EmbedGPTBigCode(vocab=96, emb_dim=32, nheads=4, nlayers=2, max_pos=32, hidden_mult=2)
                with torch.no_grad():
                    for p in m.parameters():   # non-zero biases so the "bias once" rule is exercised
                        p.normal_(0, 0.05)
                vocab = 96
            else:
                m = EmbedMixtral(LLaMAConfig(src_vocab_size=96, emb_dim=32, nheads=4, kvheads=2, nlayers=2, multiple_of=8,
                                             max_expected_seq_len=32), n_experts=4, top_k=2)
                m.reset_parameters()
                with torch.no_grad():
                    for blk in m.layers:
                        blk.moe.gate.weight.normal_(0, 0.5)
                vocab = 96
            m.eval()
            x = torch.randint(0, vocab, (2, 10))
            with torch.no_grad():
                ref_logits, ref_embeds = m(x, include_embeds=True)
                l0, cache = m(x[:, :-1], use_cache=True)
                l1, _ = m(x[:, -1:], past_key_value_states=cache, use_cache=True)
                n_before = sum(p.numel() for p in m.parameters())
                shard_for_tp(m, dist.group.WORLD)
                n_after = sum(p.numel() for p in m.parameters())
                tp_logits, tp_embeds = m(x, include_embeds=True)
                t0, tcache = m(x[:, :-1], use_cache=True)
                t1, _ = m(x[:, -1:], past_key_value_states=tcache, use_cache=True)
            res[name] = dict(logits=(tp_logits - ref_logits).abs().max().item(), embeds=(tp_embeds - ref_embeds).abs().max().item(),
                             prefill=(t0 - l0).abs().max().item(), decode=(t1 - l1).abs().max().item(),
                             scale=ref_logits.abs().max().item(), shrink=n_after / n_before)
        if rank == 0:
            torch.save(res, os.path.join(outdir, "fam.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_tp2_gpt_bigcode_and_mixtral_bases_match_unsharded():
    """The other two base-model families of the speculator registry under TP=2: multi-query attention with replicated K/V
    head and rank-0-only row-parallel biases (GPT-BigCode); expert hidden units split with a replicated router (Mixtral)."""
    outdir = tempfile.mkdtemp()
    mp.spawn(_family_worker, args=(2, free_port(), outdir), nprocs=2, join=True)
    res = torch.load(os.path.join(outdir, "fam.pt"), weights_only=False)
    for name, r in res.items():
        tol = 1e-4 * max(1.0, r["scale"])
        assert r["logits"] < tol and r["embeds"] < tol and r["prefill"] < tol and r["decode"] < tol, (name, r)
        assert r["shrink"] < 0.75, (name, r)   # the bulk of the weights really is split


@pytest.mark.parametrize("arch", ["embedllama", "embedgpt_bigcode"])   # Mixtral: TP math + init are covered in-process
def test_speculator_entrypoint_tp2_two_stages(tmp_path, arch):
    """`speculator/train_speculator.py` end to end on 2 gloo ranks: (dp, tp) = (1, 2) mesh, TP-sharded frozen base model (Llama and
    GPT-BigCode; all three families are compared with their unsharded selves above), DDP speculator on the engine, stage 1 -> stage 2 (generated continuations), final checkpoint."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(root, "speculator", "train_speculator.py"),
           f"--model_arch={arch}", "--model_variant=tiny", "--model_path=/nonexistent", "--sharding_strategy=tp",
           "--tp_size=2", "--use_dummy_dataset=True", "--num_steps=4", "--report_interval=2", "--stage2_start_step=3",
           "--stage2_batch_size=4", "--stage2_prompt_length=4", "--stage2_seq_length=8", "--n_speculator_heads=2",
           "--speculator_width=32", "--seq_length=16", "--vocab_size=512", "--batch_size=2",
           f"--ckpt_save_path={tmp_path}", f"--ckpt_load_path={tmp_path}", "--checkpoint_interval=100",
           "--comm_backend=gloo", "--use_torch_compile=False"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "step: 4" in r.stdout and "loss 2:" in r.stdout and "Checkpoint saved" in r.stdout
    assert os.path.isdir(os.path.join(tmp_path, "checkpoints", "step_4_ckp"))
