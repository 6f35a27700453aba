"""(replica x shard) rank mesh for the DP family.

  fsdp  FULL_SHARD   : replica=1,        shard=world
  hsdp  HYBRID_SHARD : replica=world/S,  shard=S   (S = hsdp_shard_size, default local device count --
                       reference behaviour, torch ``_init_utils.py:158-166``; ``--hsdp_shard_size=4``
                       gives BASELINE config "HSDP 2x4" on one 8-GPU box)
  ddp   NO_SHARD     : replica=world,    shard=1
Global rank = replica_idx * S + shard_idx (shard group = consecutive ranks = one NVSwitch domain).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch.distributed as dist


@dataclass
class DPMesh:
    world: int
    rank: int
    shard_size: int
    shard_group: Optional[object] = None      # c10d group (None when world == 1)
    replica_group: Optional[object] = None

    @property
    def replica_size(self) -> int:
        return self.world // self.shard_size

    @property
    def shard_rank(self) -> int:
        return self.rank % self.shard_size

    @property
    def replica_rank(self) -> int:
        return self.rank // self.shard_size

    def shard_group_ranks(self) -> List[int]:
        base = self.replica_rank * self.shard_size
        return list(range(base, base + self.shard_size))

    def replica_group_ranks(self) -> List[int]:
        return list(range(self.shard_rank, self.world, self.shard_size))


def resolve_shard_size(strategy: str, world: int, hsdp_shard_size: int = 0, local_world: Optional[int] = None) -> int:
    s = (strategy or "fsdp").lower()
    if s in ("ddp", "no_shard"):
        return 1
    if s in ("hsdp", "hybrid_shard"):
        size = hsdp_shard_size or local_world or world
        size = min(size, world)
        if world % size != 0:
            raise ValueError(f"world size {world} not divisible by hsdp shard size {size}")
        return size
    return world  # fsdp and -- like the reference (train_utils.py:233-234) -- anything else


def build_mesh(strategy: str, hsdp_shard_size: int = 0, local_world: Optional[int] = None) -> DPMesh:
    if not (dist.is_available() and dist.is_initialized()):
        return DPMesh(1, 0, 1)
    world, rank = dist.get_world_size(), dist.get_rank()
    S = resolve_shard_size(strategy, world, hsdp_shard_size, local_world)
    R = world // S
    shard_group = replica_group = None
    if S == world:
        shard_group = dist.group.WORLD
    elif S > 1:
        for r in range(R):
            g = dist.new_group(list(range(r * S, (r + 1) * S)))
            if rank // S == r:
                shard_group = g
    if R == world and R > 1:
        replica_group = dist.group.WORLD
    elif R > 1:
        for s in range(S):
            g = dist.new_group(list(range(s, world, S)))
            if rank % S == s:
                replica_group = g
    return DPMesh(world, rank, S, shard_group, replica_group)
