import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank); dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
from fms_fsdp_b200.models.mamba import MambaLMHeadModel, MambaConfig
from fms_fsdp_b200.utils.config_utils import get_model_config
from fms_fsdp_b200.parallel import ShardedAdamW, ShardedModel
from fms_fsdp_b200.policies import bfSixteen
c = dict(get_model_config("mamba_9.8b")); c.update(n_layer=3, attn_layer_idx=[1], vocab_size=4096)
with torch.device("meta"):
    m = MambaLMHeadModel(MambaConfig(**c))
eng = ShardedModel(m, sharding_strategy="fsdp", mixed_precision=bfSixteen, device=dev)
opt = ShardedAdamW(eng, lr=1e-3)
for st in range(3):
    x = torch.randint(0, 4096, (2, 1024), device=dev)
    loss = eng.forward_backward(x, x); gn = eng.clip_grad_norm_(1.0); opt.step()
    if rank == 0: print("step", st, loss.item(), gn.item(), [u.name for u in eng.blocks if getattr(u, "pushed", False)], flush=True)
dist.barrier(); dist.destroy_process_group()
