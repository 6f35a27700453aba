#!/bin/bash
# Regenerates the evidence under profiles/ on a B200 box (run from the repo root; every step is independent).
#   scripts/reproduce.sh [N_GPUS]          (multi-GPU steps are skipped when N_GPUS is 1)
set -u
cd "$(dirname "$0")/.."
N=${1:-1}
O=gpurun_out; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -c 'import __graft_entry__ as g; g.build(); g.smoke()'
python -m pytest tests -m gpu -x -q | tail -2
# kernels vs the ATen oracle (+ timing vs cuBLAS / cuDNN)
python scripts/gpu_diag.py gemm2 gemm elem attn ssd selscan > $O/diag_all.log 2>&1
# headline bench, both arms, per-kernel table and timeline gaps
python bench.py --steps 8 --warmup 3 --profile $O/step_kernels.txt --trace $O/trace.json > $O/bench_ours.log 2>&1
python scripts/trace_gaps.py $O/trace.json > $O/trace_gaps.txt; rm -f $O/trace.json
python bench.py --impl reference --steps 8 --warmup 3 > $O/bench_reference.log 2>&1
# ncu: one capture per hot kernel family (numbers under ncu are never bench values)
for k in gemm attn elem; do
  ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16|attn_(fwd2|bwd3)|adamw|rmsnorm|swiglu" -s 3 -c 3 \
      -o $O/prof_$k python scripts/prof_kernels.py $k > $O/ncu_$k.log 2>&1
  python scripts/ncu_summary.py $O/prof_$k.ncu-rep > $O/ncu_$k.txt
done
python scripts/sass_summary.py > $O/sass_summary.txt
python scripts/sass_listing.py > /dev/null 2>&1                    # per-kernel listings -> profiles/sass/
# collective kernels under ncu on one GPU (pointer tables over local buffers)
ncu --set full --clock-control none -k regex:'reduce_scatter|p2p_allgather|signal_barrier|scalar_allreduce' \
    -o $O/prof_comm python scripts/ncu_comm.py > $O/ncu_comm.log 2>&1 && python scripts/ncu_summary.py $O/prof_comm.ncu-rep > $O/ncu_comm.txt
# attention vs cuDNN (SDPA), instruction issue rates of the softmax building blocks, fp8 forward GEMM vs bf16
python scripts/attn_bench.py > $O/attn_bench.log 2>&1               # writes $O/attn_bench.json
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 scripts/pipe_bench.cu -o /tmp/pipe_bench && /tmp/pipe_bench > $O/pipe_issue_rates.txt
python scripts/fp8_bench.py > $O/fp8_bench.log 2>&1                 # writes $O/fp8_bench.json
python bench.py --steps 6 --warmup 3 --precision fp8 > $O/bench_fp8_nonheadline.log 2>&1
# the public entry points on the GPU: llama train + resume + HF export, mamba, speculator (TP=2 needs 2 GPUs)
bash scripts/gpu_entrypoints.sh > $O/entrypoints.log 2>&1
scripts/sanitize.sh memcheck elem gemm attn
# pipeline timeline of the attention backward, tcgen05 issue microbenchmarks
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 --use_fast_math -DB200_ATTN_TRACE -Ifms_fsdp_b200/csrc \
    scripts/attn_trace.cu -o /tmp/attn_trace -lcuda && /tmp/attn_trace 0 > $O/attn_trace.txt
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -Ifms_fsdp_b200/csrc scripts/mma_bench.cu -o /tmp/mma_bench \
    && /tmp/mma_bench > $O/mma_bench.txt
if [ "$N" -gt 1 ]; then
  $R --nproc-per-node $N --master-port 29511 scripts/gpu_multi_check.py all > $O/multi_$N.log 2>&1
  $R --nproc-per-node $N --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 --profile $O/step_kernels_$N.txt \
      > $O/bench_ours_$N.log 2>&1
  $R --nproc-per-node $N --master-port 29513 bench.py --impl reference --gpus $N --steps 6 --warmup 3 > $O/bench_reference_$N.log 2>&1
  [ "$N" -eq 8 ] && bash scripts/configs8.sh > $O/configs8.log 2>&1   # 13B selective AC, HSDP 2x4, both arms
fi
grep -h '"metric"' $O/bench_*.log | cut -c1-260
