#!/bin/bash
# On-box profiling recipe (B200_PROFILING.md): (1) per-launch device times of ~2 training steps of a 4-layer
# 7B-shaped model, (2) full ncu captures of the three hot kernels.  Outputs under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 520 --csv --log-file gpurun_out/launches.csv \
    python bench.py --nlayers 4 --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 40 -c 2 -o gpurun_out/prof_gemm \
    python scripts/gpu_diag.py gemm > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 4 -o gpurun_out/prof_attn \
    python scripts/gpu_diag.py attn > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
