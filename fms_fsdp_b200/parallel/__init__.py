from fms_fsdp_b200.parallel.engine import ShardedModel, ShardUnit  # noqa: F401
from fms_fsdp_b200.parallel.mesh import DPMesh, build_mesh  # noqa: F401
from fms_fsdp_b200.parallel.optim import ShardedAdamW  # noqa: F401
