"""MLP-speculator trainer (CLI parity with reference ``speculator/train_speculator.py:107-330``):

    torchrun --nproc_per_node=8 speculator/train_speculator.py --model_path=<hf llama dir> --model_arch=embedllama \
        --model_variant=7b --sharding_strategy=tp --tp_size=8 --n_speculator_heads=3 --stage2_start_step=15000 ...

Two-stage curriculum on top of a FROZEN base model: stage 1 trains on embeddings from one parallel forward of
the base model over ground-truth text; stage 2 on embeddings of text the base model generates itself.  The
base model is either replicated or tensor-parallel (``sharding_strategy=tp``: 2-D rank mesh (dp, tp)); the
speculator is always NO_SHARD (DDP) on the engine, i.e. its gradient path is the fused all-reduce kernel.
"""
import math
import os
import sys
import time

import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import LambdaLR

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from fms_fsdp_b200 import config  # noqa: E402
from fms_fsdp_b200.models.speculator import MLPSpeculator  # noqa: E402
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel  # noqa: E402
from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer, resolve_load_path  # noqa: E402
from fms_fsdp_b200.utils.cli import run  # noqa: E402
from fms_fsdp_b200.utils.config_utils import update_config  # noqa: E402
from fms_fsdp_b200.utils.dataloader_utils import get_data_loader, get_dummy_loader  # noqa: E402
from fms_fsdp_b200.utils.train_utils import get_profiler, setup, setup_environ_flags, torchrun_env  # noqa: E402
from speculator.train_speculator_utils import generate, get_model, train_speculator  # noqa: E402


def test_model(rank, model, arch, cfg, prompt_type="chat"):
    """Greedy 100-token smoke generation from the base model (skipped when no tokenizer ships with it).
    Reference: ``speculator/train_speculator.py:34-65``."""
    try:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(cfg.model_path)
    except Exception as e:
        if rank == 0:
            print(f"(skipping base-model text test: no tokenizer at {cfg.model_path}: {e!r})")
        return
    if prompt_type == "chat":
        prompt = ("Below is an instruction that describes a task. Write a response that appropriately completes the "
                  "request.\n\n### Instruction:\nProvide a list of instructions for preparing chicken soup.\n\n### Response:")
    else:
        prompt = ("[INST] Write code to solve the following coding problem that obeys the constraints and passes the "
                  "example test cases. Please wrap your code answer using ```:\nWrite a bubble sort function in python.\n[/INST]")
    ids = tokenizer(prompt, add_special_tokens=("llama" in arch))["input_ids"]
    dev = next(model.parameters()).device
    result = generate(model, torch.tensor(ids, dtype=torch.long, device=dev), max_new_tokens=100, use_cache=True,
                      do_sample=False, max_seq_len=8192, include_embeds=False)
    if rank == 0:
        print(f"{rank}: quick test of base model")
        print(tokenizer.decode(result.tolist()))


def get_emb_dim(model):
    """Reference: ``speculator/train_speculator.py:68-77``."""
    for k in ("emb_dim", "dim", "hidden_size"):
        if hasattr(getattr(model, "config", None), k):
            return getattr(model.config, k)
    if hasattr(model, "emb"):
        return model.emb.embedding_dim
    raise Exception("config missing embedding dimension")


def get_vocab_size(model):
    """Reference: ``speculator/train_speculator.py:80-87``."""
    for k in ("src_vocab_size", "vocab_size"):
        if hasattr(getattr(model, "config", None), k):
            return getattr(model.config, k)
    if hasattr(model, "emb"):
        return model.emb.num_embeddings
    raise Exception("config missing vocab size config")


def get_training_data_loader(rank, cfg, world_size, speculator_mesh):
    """Reference: ``speculator/train_speculator.py:90-104``."""
    if rank == 0:
        print(f"{time.time()} Constructing datasets...")
    if cfg.use_dummy_dataset:
        loader = (x for x, _ in get_dummy_loader(cfg, rank, world_size))  # unshifted sequences, like postprocess=[]
    elif cfg.sharding_strategy == "tp" and speculator_mesh is not None:
        loader = get_data_loader(cfg, speculator_mesh.get_rank(), speculator_mesh.size(), postprocess=[])
    else:
        loader = get_data_loader(cfg, rank, world_size, postprocess=[])
    if rank == 0:
        print(f"{time.time()} Datasets constructed!")
    return loader


def speculator_lr_schedule(cfg):
    """Stage 1: quadratic warm-up (min(2000, 5%)) then cosine to 0.1x; stage 2 restarts at 0.1x and anneals to 0.01x."""
    w1 = max(1, min(2000, cfg.stage2_start_step // 20))
    rem = max(1, cfg.num_steps - cfg.stage2_start_step)
    w2 = max(1, min(2000, rem // 20))
    s1 = lambda x: min(1 - (1 - min(x, w1) / w1) ** 2,  # noqa: E731
                       0.1 + 0.5 * (1 - 0.1) * (1 + math.cos(x / max(1, cfg.stage2_start_step) * math.pi)))
    s2 = lambda x: min(0.1 * (1 - (1 - min(x, w2) / w2) ** 2),  # noqa: E731
                       0.01 + 0.05 * (1 - 0.1) * (1 + math.cos(min(x, rem) / rem * math.pi)))
    return lambda x: s1(x) if x <= cfg.stage2_start_step else s2(x - cfg.stage2_start_step)


def main(**kwargs):
    """Reference: ``speculator/train_speculator.py:107-326``."""
    cfg = config.train_config()
    update_config(cfg, **kwargs)
    cfg.seq_length = cfg.seq_length + cfg.n_speculator_heads + 1

    use_cuda = torch.cuda.is_available() and cfg.comm_backend != "gloo"
    if use_cuda:
        torch.cuda.manual_seed(cfg.seed)
    torch.manual_seed(cfg.seed)
    local_rank, rank, world_size = torchrun_env()
    if rank == 0:
        print(f"{time.time()} running with these configs {cfg}")

    if world_size > 1 or "RANK" in os.environ:
        setup(cfg=cfg)
    if use_cuda:
        torch.cuda.set_device(local_rank)
        torch.cuda.empty_cache()
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    setup_environ_flags()

    base_model_mesh = speculator_mesh = None
    tp_group = None
    if cfg.sharding_strategy == "tp" and world_size > 1:
        from torch.distributed.device_mesh import init_device_mesh
        dtype_dev = "cuda" if use_cuda else "cpu"
        base_model_mesh = init_device_mesh(dtype_dev, (world_size // cfg.tp_size, cfg.tp_size), mesh_dim_names=("dp", "tp"))
        speculator_mesh = init_device_mesh(dtype_dev, (world_size,))
        tp_group = base_model_mesh["tp"].get_group()

    dtype = torch.bfloat16 if use_cuda else torch.float32
    model = get_model(cfg.model_arch, cfg.model_variant, model_path=cfg.model_path,
                      device_type="cuda" if use_cuda else "cpu", source="hf",
                      distributed_strategy=cfg.sharding_strategy, group=tp_group, dtype=dtype)
    model.eval()
    if rank == 0:
        print(f"{time.time()}", "base model loaded")
    test_model(rank, model, cfg.model_arch, cfg)

    emb_dim, vocab_size = get_emb_dim(model), get_vocab_size(model)
    speculator = MLPSpeculator(emb_dim, cfg.speculator_width, vocab_size, cfg.n_speculator_heads,
                               tie_weights=cfg.speculator_tie_weights, scale_input=cfg.speculator_scale_input)
    speculator.reset_parameters()
    if rank == 0:
        total_params = sum(p.numel() for p in speculator.parameters() if p.requires_grad)
        print(f"\n{time.time()} speculator has {total_params / 1e6} Million params\n")

    train_loader = get_training_data_loader(rank, cfg, world_size, speculator_mesh)

    from fms_fsdp_b200.policies import bfSixteen, fp32_policy
    speculator = ShardedModel(speculator, sharding_strategy="ddp", mixed_precision=bfSixteen if use_cuda else fp32_policy,
                              device=device, collective_impl=cfg.collective_impl,
                              sync_module_states=True)   # rank 0's init everywhere (reference train_speculator.py:205)
    optimizer = ShardedAdamW(speculator, lr=cfg.learning_rate, betas=(0.9, 0.95), weight_decay=0.1)

    checkpointer = Checkpointer(cfg.ckpt_save_path, 1000, "ddp", rank, local_rank)
    speculator, optimizer, train_loader, start_step, tokens_seen, is_resuming = checkpointer.load(
        speculator, optimizer, train_loader if hasattr(train_loader, "dataset") else None,
        path=resolve_load_path(cfg.ckpt_load_path), is_compiled=cfg.use_torch_compile)
    if train_loader is None:
        train_loader = get_training_data_loader(rank, cfg, world_size, speculator_mesh)
    if not is_resuming:
        start_step = 0
        for g in optimizer.param_groups:
            g["initial_lr"] = cfg.learning_rate

    schedule = speculator_lr_schedule(cfg)
    scheduler = LambdaLR(optimizer, lambda x: schedule(x + start_step))
    profiler = get_profiler(cfg, rank)

    if rank == 0:
        print(f"{time.time()} Training for {cfg.num_steps} steps")
    torch.cuda.empty_cache() if use_cuda else None
    train_speculator(cfg, model, speculator, local_rank, rank, train_loader, optimizer, scheduler, checkpointer,
                     start_step, tokens_seen, profiler, base_model_mesh)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    run(main)
