"""Sharded runtime vs a plain single-process oracle (BASELINE config #1: Llama2-tiny, world 2, CPU/gloo).

Multi-rank runs are real processes over gloo on 127.0.0.1; the oracle is the unsharded model trained
with torch.optim.AdamW + clip_grad_norm_ on the concatenated global batch."""
import copy
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import free_port
from fms_fsdp_b200.models.llama import LLaMA, LLaMABlock
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel
from fms_fsdp_b200.parallel.layout import build_layout, dim0_chunk
from fms_fsdp_b200.parallel.mesh import resolve_shard_size
from fms_fsdp_b200.policies import apply_fsdp_checkpointing, bfSixteen, fp32_policy
from fms_fsdp_b200.utils.config_utils import get_model_config

STEPS, B, S, V = 3, 2, 32, 1024


def _batch(rank, step):
    g = torch.Generator().manual_seed(1000 * step + rank)
    return torch.randint(0, V, (B, S), generator=g)


def _oracle(world, steps=STEPS, lr=1e-3):
    torch.manual_seed(0)
    m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
    opt = torch.optim.AdamW(m.parameters(), lr=lr, betas=(0.9, 0.95), weight_decay=0.1)
    losses, norms = [], []
    for st in range(steps):
        opt.zero_grad()
        tot = 0.0
        for r in range(world):
            x = _batch(r, st)
            l = m(x, labels=x) / world
            l.backward()
            tot += l.item()
        norms.append(torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0).item())
        opt.step()
        losses.append(tot)
    return losses, norms, {k: v.detach().clone() for k, v in m.state_dict().items()}


def _worker(rank, world, port, strategy, shard_size, ac, outdir, ckpt_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
        if ac:
            apply_fsdp_checkpointing(m, LLaMABlock, ac)
        eng = ShardedModel(m, sharding_strategy=strategy, hsdp_shard_size=shard_size, mixed_precision=fp32_policy,
                           device="cpu", collective_impl="torch")
        opt = ShardedAdamW(eng, lr=1e-3)
        losses, norms = [], []
        for st in range(STEPS):
            x = _batch(rank, st)
            loss = eng.forward_backward(x, x)
            norms.append(eng.clip_grad_norm_(1.0).item())
            opt.step()
            t = loss.clone(); dist.all_reduce(t); losses.append(t.item() / world)
        sd = eng.full_state_dict()
        if ckpt_dir:
            from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
            Checkpointer(ckpt_dir, 5, strategy, rank, rank).save(STEPS, eng, opt, None, tokens_seen=123)
        if rank == 0:
            torch.save(dict(losses=losses, norms=norms, sd=sd), os.path.join(outdir, "out.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, strategy, shard_size=0, ac=None, ckpt_dir=None):
    outdir = tempfile.mkdtemp()
    mp.spawn(_worker, args=(world, free_port(), strategy, shard_size, ac, outdir, ckpt_dir), nprocs=world, join=True)
    return torch.load(os.path.join(outdir, "out.pt"), weights_only=False)


def _check(out, world):
    losses, norms, sd = _oracle(world)
    assert out["losses"] == pytest.approx(losses, rel=2e-4, abs=2e-4)
    assert out["norms"] == pytest.approx(norms, rel=2e-3)
    for k, v in sd.items():
        assert torch.allclose(out["sd"][k], v, atol=2e-5, rtol=1e-4), k


@pytest.mark.parametrize("strategy,shard,ac", [("fsdp", 0, None), ("fsdp", 0, "1/2"), ("ddp", 0, None)])
def test_world2_matches_oracle(strategy, shard, ac):
    _check(_run(2, strategy, shard, ac), 2)


def test_world2_poisoned_release_still_matches_oracle(monkeypatch):
    """Debug trap of SURVEY 5.2: gathered parameters are NaN-filled the moment a unit is released; the schedule
    (incl. selective recompute) must never read them afterwards, so the run still equals the oracle."""
    monkeypatch.setenv("FMS_B200_POISON", "1")
    _check(_run(2, "fsdp", 0, "1/2"), 2)


def test_hsdp_2x2_matches_oracle():
    _check(_run(4, "hsdp", 2), 4)


def test_single_process_bf16_policy_runs_and_learns(tiny_llama):
    eng = ShardedModel(tiny_llama, mixed_precision=bfSixteen, device="cpu")
    opt = ShardedAdamW(eng, lr=3e-3)
    x = _batch(0, 0)
    l0 = None
    for _ in range(8):
        l = eng.forward_backward(x, x); eng.clip_grad_norm_(1.0); opt.step()
        l0 = l0 if l0 is not None else l.item()
    assert l.item() < l0


def test_reference_style_loop_matches_fast_path(tiny_llama):
    """``out = model(x); loss = CE(out); loss.backward()`` goes through the same schedule."""
    a, b = copy.deepcopy(tiny_llama), copy.deepcopy(tiny_llama)
    ea, eb = ShardedModel(a, device="cpu"), ShardedModel(b, device="cpu")
    x = _batch(0, 0)
    la = ea.forward_backward(x, x); na = ea.clip_grad_norm_(1.0)
    out = eb(x)
    lb = torch.nn.functional.cross_entropy(out.view(-1, out.size(-1)), x.view(-1))
    lb.backward(); nb = eb.clip_grad_norm_(1.0)
    assert la.item() == pytest.approx(lb.item(), rel=1e-5)
    assert na.item() == pytest.approx(nb.item(), rel=1e-4)


def test_checkpoint_reshard_2_to_1_and_resume():
    from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
    ck = tempfile.mkdtemp()
    out = _run(2, "fsdp", ckpt_dir=ck)
    torch.manual_seed(1)
    m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
    eng = ShardedModel(m, device="cpu"); opt = ShardedAdamW(eng, lr=1e-3)
    _, _, _, step, ntok, resuming = Checkpointer(ck, 5, "fsdp", 0, 0).load(eng, opt, None, path="")
    assert (step, ntok, resuming) == (STEPS, 123, True) and opt._step == STEPS
    sd = eng.full_state_dict()
    for k, v in out["sd"].items():
        assert torch.equal(sd[k], v), k
    # exporter-style read: plain DCP load into full tensors with the reference key layout
    import torch.distributed.checkpoint as dcp
    full = {"model_state": {k: torch.empty_like(v) for k, v in out["sd"].items()}}
    dcp.load(full, checkpoint_id=os.path.join(ck, "checkpoints", f"step_{STEPS}_ckp"), no_dist=True)
    for k, v in out["sd"].items():
        assert torch.equal(full["model_state"][k], v), k


def test_layout_math():
    lay = build_layout("u", [("a", (5, 7)), ("b", (3,)), ("c", (64, 2))], 4)
    assert lay.total % (4 * 64) == 0 and [s.offset % 64 for s in lay.slots] == [0, 0, 0]
    covered = sum(lay.overlap(s, r)[2] for s in lay.slots for r in range(4))
    assert covered == lay.used == 35 + 3 + 128
    assert [dim0_chunk(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert resolve_shard_size("hsdp", 8, 4) == 4 and resolve_shard_size("ddp", 8) == 1
    assert resolve_shard_size("whatever", 8) == 8  # unknown -> full shard, like the reference


def _sync_worker(rank, world, port, outdir, mismatch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                     # deliberately different initial weights per rank
        cfg = get_model_config("llama2_tiny")
        if mismatch == "depth" and rank == 1:
            cfg.nlayers += 1                              # a model that differs across ranks must be rejected
        if mismatch == "width" and rank == 1:
            cfg.hidden_grow_factor = cfg.hidden_grow_factor * 2
        m = LLaMA(cfg); m.reset_parameters()
        err = None
        try:
            eng = ShardedModel(m, sharding_strategy="fsdp", mixed_precision=fp32_policy, device="cpu",
                               collective_impl="torch", sync_module_states=True)
            sd = eng.full_state_dict()
        except RuntimeError as ex:
            err, sd = str(ex), None
        if rank == 0:
            torch.manual_seed(100)
            ref = LLaMA(get_model_config("llama2_tiny")); ref.reset_parameters()
            torch.save(dict(err=err, sd=sd, ref=ref.state_dict()), os.path.join(outdir, "out.pt"))
    finally:
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()


def test_sync_module_states_broadcasts_rank0_init():
    outdir = tempfile.mkdtemp()
    mp.spawn(_sync_worker, args=(2, free_port(), outdir, False), nprocs=2, join=True)
    out = torch.load(os.path.join(outdir, "out.pt"), weights_only=False)
    assert out["err"] is None
    for k, v in out["ref"].items():
        assert torch.equal(out["sd"][k], v), k


@pytest.mark.parametrize("kind", ["depth", "width"])
def test_models_that_differ_across_ranks_are_rejected(kind):
    outdir = tempfile.mkdtemp()
    mp.spawn(_sync_worker, args=(2, free_port(), outdir, kind), nprocs=2, join=True)
    out = torch.load(os.path.join(outdir, "out.pt"), weights_only=False)
    assert out["err"] is not None and "differs across ranks" in out["err"]


def test_wrapping_policy_decides_the_shard_units():
    """``get_wrapper(block)`` is consumed by the engine (reference ``policies/wrapping.py:6-14`` contract): blocks the
    predicate accepts become units of their own, the others stay resident in the root unit -- same training result."""
    from fms_fsdp_b200.policies import get_wrapper
    ref_losses, ref_norms, _ = _oracle(1)

    def run(policy):
        torch.manual_seed(0)
        m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
        eng = ShardedModel(m, sharding_strategy="fsdp", mixed_precision=fp32_policy, device="cpu",
                           collective_impl="torch", auto_wrap_policy=policy)
        opt = ShardedAdamW(eng, lr=1e-3)
        out = []
        for st in range(STEPS):
            x = _batch(0, st)
            loss = eng.forward_backward(x, x)
            out.append((loss.item(), eng.clip_grad_norm_(1.0).item()))
            opt.step()
        return eng, out

    eng, out = run(get_wrapper(LLaMABlock))
    n_layers = get_model_config("llama2_tiny").nlayers
    assert len(eng.blocks) == n_layers and all(u is not None for _, u in eng._chain)
    layers = None

    def only_first(module):            # a policy that wraps just the first block
        return module is layers[0]
    torch.manual_seed(0)
    m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
    layers = list(m.layers)
    eng2 = ShardedModel(m, sharding_strategy="fsdp", mixed_precision=fp32_policy, device="cpu", collective_impl="torch",
                        auto_wrap_policy=only_first)
    assert len(eng2.blocks) == 1 and [u is not None for _, u in eng2._chain] == [True] + [False] * (n_layers - 1)
    root_names = {n for n, _ in eng2.root.params}
    assert any(n.startswith("layers.1.") for n in root_names) and not any(n.startswith("layers.0.") for n in root_names)
    opt2 = ShardedAdamW(eng2, lr=1e-3)
    for st in range(STEPS):
        x = _batch(0, st)
        loss = eng2.forward_backward(x, x)
        gn = eng2.clip_grad_norm_(1.0).item()
        opt2.step()
        assert abs(loss.item() - out[st][0]) < 1e-5 and abs(gn - out[st][1]) < 1e-4
        assert abs(loss.item() - ref_losses[st]) < 1e-4


def _reload_worker(rank, world, port, ckpt_dir, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
        torch.manual_seed(7)                                   # different weights than the checkpoint's
        m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
        eng = ShardedModel(m, sharding_strategy="fsdp", mixed_precision=fp32_policy, device="cpu", collective_impl="torch")
        opt = ShardedAdamW(eng, lr=1e-3)
        _, _, _, step, ntok, resuming = Checkpointer(ckpt_dir, 5, "fsdp", rank, rank).load(eng, opt, None, path="")
        sd = eng.full_state_dict()
        x = _batch(rank, STEPS)
        loss = eng.forward_backward(x, x); gn = eng.clip_grad_norm_(1.0); opt.step()     # and training continues
        if rank == 0:
            torch.save(dict(sd=sd, step=step, ntok=ntok, resuming=resuming, opt_step=opt._step, loss=loss.item(),
                            gn=gn.item()), os.path.join(outdir, "reload.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("save_world,load_world", [(2, 4), (4, 2)])
def test_checkpoint_reshards_across_world_sizes(save_world, load_world):
    """A sharded checkpoint written by W ranks loads into a job of a different size (the DCP chunks are per-parameter
    dim-0 shards of the CURRENT world size; reference behaviour: resharding-on-load), model + optimizer + step."""
    ck = tempfile.mkdtemp()
    out = _run(save_world, "fsdp", ckpt_dir=ck)
    outdir = tempfile.mkdtemp()
    mp.spawn(_reload_worker, args=(load_world, free_port(), ck, outdir), nprocs=load_world, join=True)
    r = torch.load(os.path.join(outdir, "reload.pt"), weights_only=False)
    assert (r["step"], r["ntok"], r["resuming"], r["opt_step"]) == (STEPS, 123, True, STEPS + 1)
    for k, v in out["sd"].items():
        assert torch.equal(r["sd"][k], v), k
    assert r["loss"] == r["loss"] and r["gn"] > 0


def test_collective_plan_single_and_multi_node():
    """Which collectives run where (``parallel/comm.plan_collectives``): NVLink peer kernels need a group inside one node."""
    import pytest as _pt
    from fms_fsdp_b200.parallel.comm import plan_collectives as plan
    # one node: everything fused, any strategy
    assert plan("auto", "cuda", 8, 8, 8) == ("fused", "fused")            # fsdp
    assert plan("auto", "cuda", 8, 4, 8) == ("fused", "fused")            # hsdp 2x4 on one box
    assert plan("auto", "cuda", 8, 1, 8) == ("fused", "fused")            # ddp
    assert plan("fused", "cuda", 2, 2, 2) == ("fused", "fused")
    assert plan("torch", "cuda", 8, 8, 8)[0] == "torch"
    assert plan("auto", "cuda", 1, 1, 1)[0] == "torch" and plan("fused", "cuda", 1, 1, 8)[0] == "torch"
    assert plan("auto", "cpu", 4, 4, 4)[0] == "torch"
    with _pt.raises(ValueError):
        plan("fused", "cpu", 2, 2, 2)
    # the reference's production layout: 16 nodes x 8 GPUs, shard inside the node, replicate across nodes
    assert plan("auto", "cuda", 128, 8, 8) == ("fused", "nccl")
    assert plan("auto", "cuda", 16, 4, 8) == ("fused", "nccl")            # two shard groups per node
    assert plan("fused", "cuda", 16, 8, 8) == ("fused", "nccl")
    # shards spanning nodes (multi-node FSDP), multi-node DDP, shard size not dividing the node
    assert plan("auto", "cuda", 16, 16, 8)[0] == "torch"
    assert plan("auto", "cuda", 16, 1, 8)[0] == "torch"
    assert plan("auto", "cuda", 24, 3, 8)[0] == "torch"
    with _pt.raises(ValueError, match="do not fit inside a node"):
        plan("fused", "cuda", 16, 16, 8)
    # more ranks in a single-node group than a signal pad holds
    assert plan("auto", "cuda", 64, 64, 64)[0] == "torch"


@pytest.mark.parametrize("compiled", [False, True])
def test_checkpoint_in_the_torch_fsdp_key_layout_loads_and_training_continues(compiled):
    """A checkpoint laid out the way the reference writes it -- torch DCP, ``model_state.<fqn>`` (``model_state._orig_mod.<fqn>``
    under torch.compile), ``optimizer_state.state.<fqn>.{exp_avg,exp_avg_sq,step}``, ``optimizer_state.param_groups``,
    ``metadata.pth`` (reference ``checkpointing_utils.py:283-310``) -- produced here from a PLAIN torch model and
    ``torch.optim.AdamW`` (torch FSDP itself needs an accelerator), must resume in this runtime: the next optimizer step equals
    the plain model's next step."""
    import torch.distributed.checkpoint as dcp
    from torch.distributed.checkpoint import FileSystemWriter
    from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
    torch.manual_seed(0)
    plain = LLaMA(get_model_config("llama2_tiny")); plain.reset_parameters()
    opt = torch.optim.AdamW(plain.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1)

    def plain_step(st):
        opt.zero_grad()
        x = _batch(0, st)
        plain(x, labels=x).backward()
        torch.nn.utils.clip_grad_norm_(plain.parameters(), 1.0)
        opt.step()

    for st in range(2):
        plain_step(st)
    names = {p: n for n, p in plain.named_parameters()}
    pre = "_orig_mod." if compiled else ""
    osd = opt.state_dict()
    fqn_of = {i: names[p] for i, p in enumerate(opt.param_groups[0]["params"])}
    state = {pre + fqn_of[i]: {k: v.clone() if torch.is_tensor(v) else v for k, v in s.items()} for i, s in osd["state"].items()}
    groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": [pre + fqn_of[i] for i in g["params"]]}
              for g in osd["param_groups"]]
    ckdir = tempfile.mkdtemp()
    save_name = os.path.join(ckdir, "checkpoints", "step_2_ckp")
    os.makedirs(save_name)
    dcp.save({"model_state": {pre + k: v.clone() for k, v in plain.state_dict().items()},
              "optimizer_state": {"state": state, "param_groups": groups}},
             storage_writer=FileSystemWriter(save_name, single_file_per_rank=True), no_dist=True)
    torch.save({"step": 2, "tokens_seen": 77}, os.path.join(save_name, "metadata.pth"))

    torch.manual_seed(5)     # different init: everything must come from the checkpoint
    m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
    eng = ShardedModel(m, device="cpu"); ours = ShardedAdamW(eng, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1)
    _, _, _, step, ntok, resuming = Checkpointer(ckdir, 5, "fsdp", 0, 0).load(eng, ours, None, path="", is_compiled=compiled)
    assert (step, ntok, resuming) == (2, 77, True) and ours._step == 2
    x = _batch(0, 2)
    eng.forward_backward(x, x); eng.clip_grad_norm_(1.0); ours.step()
    plain_step(2)
    mine = eng.full_state_dict()
    for k, v in plain.state_dict().items():
        assert torch.allclose(mine[k], v, atol=2e-6, rtol=1e-5), (k, (mine[k] - v).abs().max())


@pytest.mark.parametrize("strategy,world,shard", [("hsdp", 4, 2), ("ddp", 2, 0)])
def test_hsdp_and_ddp_checkpoints_load_into_other_layouts(strategy, world, shard):
    """HSDP: only the shard group of replica 0 writes the tensor files (reference ``_do_save``); DDP: rank 0 writes.  Either
    checkpoint must come back -- model, AdamW moments, step -- in a single-process job and in an FSDP job of 2 ranks."""
    from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
    ck = tempfile.mkdtemp()
    out = _run(world, strategy, shard, ckpt_dir=ck)
    step_dir = os.path.join(ck, "checkpoints", f"step_{STEPS}_ckp")
    n_files = len([f for f in os.listdir(step_dir) if f.endswith(".distcp")])
    assert n_files == (2 if strategy == "hsdp" else 1), os.listdir(step_dir)

    torch.manual_seed(3)
    m = LLaMA(get_model_config("llama2_tiny")); m.reset_parameters()
    eng = ShardedModel(m, device="cpu"); opt = ShardedAdamW(eng, lr=1e-3)
    _, _, _, step, ntok, resuming = Checkpointer(ck, 5, "fsdp", 0, 0).load(eng, opt, None, path="")
    assert (step, ntok, resuming, opt._step) == (STEPS, 123, True, STEPS)
    sd = eng.full_state_dict()
    for k, v in out["sd"].items():
        assert torch.equal(sd[k], v), k
    assert all(float(u.exp_avg.abs().sum()) > 0 and float(u.exp_avg_sq.sum()) > 0 for u in eng.units)

    if strategy == "ddp":
        return                      # the sharded reload below is exercised by the hsdp case
    outdir = tempfile.mkdtemp()
    mp.spawn(_reload_worker, args=(2, free_port(), ck, outdir), nprocs=2, join=True)
    r = torch.load(os.path.join(outdir, "reload.pt"), weights_only=False)
    assert (r["step"], r["ntok"], r["resuming"], r["opt_step"]) == (STEPS, 123, True, STEPS + 1)
    for k, v in out["sd"].items():
        assert torch.equal(r["sd"][k], v), k


def test_scaled_losses_scale_the_gradients(tiny_llama):
    """``(loss * k).backward()`` -- gradient accumulation, loss weighting -- through both public forms: logits + external
    cross-entropy, and the fused ``model(tokens, labels)`` loss."""
    x = _batch(0, 0)
    norms = []
    for scale, fused in ((1.0, False), (0.25, False), (0.25, True)):
        eng = ShardedModel(copy.deepcopy(tiny_llama), device="cpu")
        if fused:
            loss = eng(x, x)
        else:
            out = eng(x)
            loss = torch.nn.functional.cross_entropy(out.view(-1, out.size(-1)), x.view(-1))
        (loss * scale).backward()
        norms.append(eng.clip_grad_norm_(1e9).item())
    assert norms[1] == pytest.approx(0.25 * norms[0], rel=1e-5) and norms[2] == pytest.approx(0.25 * norms[0], rel=1e-4)
