"""Small services of the training loop that the entry-point drills only touch from subprocesses: LR schedule, profiler and
tracker hooks, MFU accounting, optimizer state round trip, engine introspection, mesh / layout helpers."""
import math
import sys
import types

import pytest
import torch

from fms_fsdp_b200.config import train_config
from fms_fsdp_b200.models.llama import LLaMA
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel
from fms_fsdp_b200.parallel.layout import build_layout
from fms_fsdp_b200.parallel.mesh import DPMesh
from fms_fsdp_b200.utils import train_utils as TU
from fms_fsdp_b200.utils.config_utils import get_model_config


def test_lr_schedule_is_the_reference_formula():
    """Reference ``main_training_llama.py:137-148``: quadratic warm-up over min(2000, steps/20), cosine to 10 %; 'annealing' is
    linear to zero."""
    cfg = train_config()
    cfg.num_steps, cfg.training_stage = 100000, "initial"
    f = TU.lr_schedule_fn(cfg)
    warm = 2000
    for x in (0, 1, 500, 1999, 2000, 2001, 50000, 99999, 100000, 150000):
        want = min(1 - (1 - min(x, warm) / warm) ** 2,
                   0.1 + 0.5 * (1 - 0.1) * (1 + math.cos(min(x, cfg.num_steps) / cfg.num_steps * math.pi)))
        assert f(x) == pytest.approx(want, abs=1e-12)
    assert f(0) == 0 and f(2000) == pytest.approx(1.0, abs=2e-3) and f(100000) == pytest.approx(0.1) and f(10 ** 7) == pytest.approx(0.1)
    cfg.num_steps = 1000                       # short runs: warm-up is steps / 20
    g = TU.lr_schedule_fn(cfg)
    assert g(50) == pytest.approx(min(1.0, 0.1 + 0.45 * (1 + math.cos(0.05 * math.pi)))) and g(25) == pytest.approx(0.75)
    cfg.training_stage = "annealing"
    h = TU.lr_schedule_fn(cfg)
    assert h(0) == 1 and h(250) == pytest.approx(0.75) and h(1000) == 0


def test_profiler_hook_respects_flags():
    cfg = train_config()
    cfg.use_profiler = False
    assert TU.get_profiler(cfg, 0) is None
    cfg.use_profiler, cfg.profiler_rank0_only = True, True
    assert TU.get_profiler(cfg, 1) is None
    p = TU.get_profiler(cfg, 0)
    assert isinstance(p, torch.profiler.profile)
    cfg.profiler_rank0_only = False
    assert isinstance(TU.get_profiler(cfg, 3), torch.profiler.profile)


def test_tracker_hooks(monkeypatch):
    cfg = train_config()
    cfg.tracker = None
    assert TU._init_tracker(cfg, 0) is None
    cfg.tracker = "tensorboard"
    with pytest.raises(ValueError, match="not supported"):
        TU._init_tracker(cfg, 0)
    for name in ("wandb", "aim"):                      # neither is installed in this image
        monkeypatch.setitem(sys.modules, name, None)
        cfg.tracker = name
        with pytest.raises(ImportError, match=f"{name} is not installed"):
            TU._init_tracker(cfg, 0)
    # a stand-in wandb: rank 0 gets the log function, other ranks nothing
    calls = {}
    fake = types.SimpleNamespace(init=lambda **kw: calls.update(kw), log=lambda *a, **k: None,
                                 errors=types.SimpleNamespace(UsageError=RuntimeError))
    monkeypatch.setitem(sys.modules, "wandb", fake)
    cfg.tracker, cfg.tracker_project_name, cfg.tracker_run_id = "wandb", "proj", "abc"
    assert TU._init_tracker(cfg, 1) is None and not calls
    assert TU._init_tracker(cfg, 0) is fake.log and calls["project"] == "proj" and calls["id"] == "abc"
    assert fake.config["model_variant"] == cfg.model_variant


def test_flops_accounting_and_peak():
    c = get_model_config("llama2_7b")
    n = 6738415616
    per_tok = TU.model_flops_per_token(n, c.nlayers, c.emb_dim, 4096)
    assert per_tok == 6.0 * n + 12.0 * 32 * 4096 * 4096
    assert 500 < TU.peak_tflops() < 3000          # measured cuBLAS figure when the driver wrote one, else the recipe's fallback


def test_optimizer_state_dict_round_trip_continues_identically(tiny_llama):
    import copy
    a, b = copy.deepcopy(tiny_llama), copy.deepcopy(tiny_llama)
    ea, eb = ShardedModel(a, device="cpu"), ShardedModel(b, device="cpu")
    oa, ob = ShardedAdamW(ea, lr=1e-3), ShardedAdamW(eb, lr=5e-4)
    x = torch.randint(0, 1024, (2, 32))
    for _ in range(2):
        ea.forward_backward(x, x); ea.clip_grad_norm_(1.0); oa.step()
    sd = oa.state_dict()
    assert sd["step"] == 2 and set(sd["units"]) == {u.name for u in ea.units}
    eb.load_full_state_dict(ea.full_state_dict())
    ob.load_state_dict(sd)
    assert ob._step == 2 and ob.param_groups[0]["lr"] == 1e-3
    for e, o in ((ea, oa), (eb, ob)):
        e.forward_backward(x, x); e.clip_grad_norm_(1.0); o.step()
    fa, fb = ea.full_state_dict(), eb.full_state_dict()
    for k in fa:
        assert torch.equal(fa[k], fb[k]), k


def test_engine_introspection(tiny_llama):
    n = sum(p.numel() for p in tiny_llama.parameters())
    eng = ShardedModel(tiny_llama, device="cpu")
    assert eng.param_count() == n
    r = eng.extra_repr()
    assert "collectives=torch" in r and "units" in r
    with eng.summon_full_params():
        full = {k: v.clone() for k, v in eng.module.state_dict().items()}
    ref = eng.full_state_dict()
    assert set(full) == set(ref) and all(torch.equal(full[k].float(), ref[k].float()) for k in ref)


def test_mesh_and_layout_helpers():
    m = DPMesh(world=8, rank=5, shard_size=4)
    assert (m.replica_size, m.shard_rank, m.replica_rank) == (2, 1, 1)
    assert m.shard_group_ranks() == [4, 5, 6, 7] and m.replica_group_ranks() == [1, 5]
    lay = build_layout("u", [("n", (7,)), ("w", (64, 24)), ("b", (5,))], 4)
    assert lay.signature() == build_layout("v", [("n", (7,)), ("w", (64, 24)), ("b", (5,))], 4).signature()
    assert lay.signature() != build_layout("u", [("n", (7,)), ("w", (64, 32)), ("b", (5,))], 4).signature()
    covered = sum(s.numel for s in lay.slots) + sum(hi - lo for lo, hi in lay.gaps())
    assert covered == lay.total                      # slots + alignment gaps tile the flat buffer exactly


def test_param_init_function_materialises_meta_modules():
    from fms_fsdp_b200.policies import param_init_function
    with torch.device("meta"):
        m = LLaMA(get_model_config("llama2_tiny"))
    blk = m.layers[0]
    assert next(blk.parameters()).is_meta
    param_init_function(blk)                                    # like torch FSDP's param_init_fn: one module, no recursion
    assert next(blk.parameters()).is_meta
    for sub in reversed(list(blk.modules())):                   # the runtime applies it children-first over the unit
        param_init_function(sub, torch.device("cpu"))
    assert not any(p.is_meta for p in blk.parameters()) and all(torch.isfinite(p).all() for p in blk.parameters())
    assert float(blk.ln.weight.min()) == 1.0 and 0 < float(blk.attn.dense.weight.std()) < 0.05


def test_resolve_load_path(tmp_path):
    from fms_fsdp_b200.utils.checkpointing_utils import resolve_load_path
    run = tmp_path / "run"
    ck = run / "checkpoints" / "step_5_ckp"
    ck.mkdir(parents=True)
    (ck / "metadata.pth").write_bytes(b"x")
    f = tmp_path / "model.pth"
    f.write_bytes(b"x")
    assert resolve_load_path(str(f)) == str(f)                                              # single file: as is
    assert resolve_load_path(str(run)) == str(run) + "/checkpoints/"                        # run directory: reference rule
    assert resolve_load_path(str(ck)) == str(ck)                                            # one checkpoint folder
    assert resolve_load_path(str(run / "checkpoints")) == str(run / "checkpoints")          # folder of checkpoints
    assert resolve_load_path(str(tmp_path / "nothing")) == str(tmp_path / "nothing") + "/checkpoints/"


def test_doctor_reports_the_environment_without_a_gpu():
    from fms_fsdp_b200.utils import doctor
    info = doctor.collect()
    assert info["torch"] == torch.__version__ and isinstance(info["gpus"], list)
    if not torch.cuda.is_available():
        assert info["gpus"] == [] and info["problems"] == [] and any("ATen path" in n for n in info["notes"])
        assert doctor.main() == 0


def test_scheduler_driven_learning_rate_matches_torch_adamw(tiny_llama):
    """``LambdaLR`` drives ``ShardedAdamW.param_groups[0]["lr"]`` exactly as it drives ``torch.optim.AdamW``: warm-up + decay
    over several steps with gradient clipping -- parameters stay equal to the plain-PyTorch run."""
    import copy
    from torch.optim.lr_scheduler import LambdaLR
    ref = copy.deepcopy(tiny_llama)
    eng = ShardedModel(tiny_llama, device="cpu")
    sched_fn = lambda s: min(1.0, (s + 1) / 3) * (0.5 ** (s // 4))     # noqa: E731
    ours, theirs = ShardedAdamW(eng, lr=2e-3), torch.optim.AdamW(ref.parameters(), lr=2e-3, betas=(0.9, 0.95), weight_decay=0.1)
    so, st = LambdaLR(ours, sched_fn), LambdaLR(theirs, sched_fn)
    g = torch.Generator().manual_seed(0)
    for _ in range(6):
        x = torch.randint(0, 1024, (2, 24), generator=g)
        eng.forward_backward(x, x); eng.clip_grad_norm_(0.5); ours.step(); so.step()
        theirs.zero_grad(); ref(x, labels=x).backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5); theirs.step(); st.step()
        assert so.get_last_lr() == st.get_last_lr()
    mine = eng.full_state_dict()
    for k, v in ref.state_dict().items():
        assert torch.allclose(mine[k], v, atol=3e-6, rtol=1e-5), (k, (mine[k] - v).abs().max())


def test_train_loop_reports_to_the_tracker_with_the_reference_keys(tiny_llama, monkeypatch, tmp_path, capsys):
    """``train()`` in-process on CPU with a stand-in ``wandb``: one tracker call per report step carrying exactly the keys the
    reference logs (``train_utils.py:151-165``), the reference's stdout lines, and a final checkpoint."""
    from torch.optim.lr_scheduler import LambdaLR
    from fms_fsdp_b200.utils.checkpointing_utils import Checkpointer
    from fms_fsdp_b200.utils.dataloader_utils import get_dummy_loader
    logged = []
    fake = types.SimpleNamespace(init=lambda **kw: None, log=lambda vals, step=None: logged.append((step, dict(vals))),
                                 errors=types.SimpleNamespace(UsageError=RuntimeError))
    monkeypatch.setitem(sys.modules, "wandb", fake)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    cfg = train_config()
    cfg.model_variant, cfg.tracker, cfg.use_dummy_dataset = "llama2_tiny", "wandb", True
    cfg.seq_length, cfg.batch_size, cfg.vocab_size, cfg.num_steps = 32, 2, 1024, 3
    cfg.report_interval, cfg.checkpoint_interval, cfg.ckpt_save_path = 1, 100, str(tmp_path)
    eng = ShardedModel(tiny_llama, device="cpu")
    opt = ShardedAdamW(eng, lr=1e-3)
    TU.train(cfg, eng, 0, 0, get_dummy_loader(cfg, 0, 1), opt, LambdaLR(opt, lambda s: 1.0), None,
             Checkpointer(str(tmp_path), 2, "fsdp", 0, 0), 0, 0)
    assert [s for s, _ in logged] == [1, 2, 3]
    assert set(logged[0][1]) == {"learning rate", "loss", "gradient norm", "token seen",
                                 "current throughput (token per gpu per sec)", "overall throughput (token per gpu per sec)",
                                 "gpu reserved memory", "gpu allocated memory"}
    assert logged[2][1]["token seen"] == 3 * 2 * 32 and all(math.isfinite(v["loss"]) for _, v in logged)
    out = capsys.readouterr().out
    for line in ("step: 3", "loss:", "LR:", "tokens seen: 192", "gradient norm:", "overall token per day:", "Checkpoint saved"):
        assert line in out, line
