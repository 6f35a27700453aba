"""Collective kernels under ncu on ONE GPU (pointer tables over local buffers, 7B-block-sized): launched by
  ncu --set full --clock-control none -k regex:'reduce_scatter|p2p_allgather|signal_barrier|scalar_allreduce' -o gpurun_out/prof_comm python scripts/ncu_comm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fms_fsdp_b200.ops import cuda_kernels as CK
C = CK._C
dev = "cuda"
W = 8
n = 202383360 // (W * 64) * 64            # one rank's shard of a Llama2-7B block
table = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
slots = torch.randn(W * n, device=dev).bfloat16()            # staging buffer [src rank][shard]
tab = torch.tensor([slots.data_ptr() + s * n * 2 for s in range(W)], dtype=torch.int64, device=dev)
out = torch.empty(n, device=dev); ss = torch.zeros((), device=dev)
for _ in range(2):
    C.reduce_scatter(tab, out, 0, W, 0, True, 1.0 / W, ss)        # the local slot sum of the push path
shards = [torch.randn(n, device=dev).bfloat16() for _ in range(W)]
full = torch.empty(W * n, dtype=torch.bfloat16, device=dev)
for _ in range(2):
    C.p2p_allgather(table(shards), full, n * 2, W, 0)
pad = torch.zeros(32 * 16, dtype=torch.int32, device=dev); anchor = torch.zeros(1, device=dev)
C.signal_barrier(table([pad]), 1, 0, 1, anchor, 0, 0)
buf = torch.zeros(1024, dtype=torch.int32, device=dev); v = torch.ones(1, device=dev)
C.scalar_allreduce(table([buf]), 1, 0, 1, v)
torch.cuda.synchronize()
