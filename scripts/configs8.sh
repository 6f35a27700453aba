#!/bin/bash
# BASELINE.json configurations 2-5 on N GPUs in one call (each is one bench.py JSON line, written to gpurun_out/cfg_*.log):
#   7B FSDP (headline), 13B FSDP + selective AC 1/2 (ours and the reference arm), 7B HSDP 2 x N/2, mamba_9.8b FSDP
N=${1:-8}
O=gpurun_out; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; timeout 420 $T bench.py --gpus $N "$@" > $O/cfg_${N}_$name.log 2>&1; echo "== $name rc=$?"; grep -o '{"metric.*\|{"impl.*' $O/cfg_${N}_$name.log | cut -c1-400; }
run 7b_fsdp --steps ${STEPS:-10} --warmup 4
run 13b_ac_half --model llama2_13b --ac 1/2 --steps 6 --warmup 3
run 7b_hsdp --sharding hsdp --hsdp_shard_size $((N / 2)) --steps 6 --warmup 3
run mamba_9.8b --model mamba_9.8b --steps 4 --warmup 3
if [ "${REF13:-1}" = "1" ]; then run 13b_ac_half_reference --impl reference --model llama2_13b --ac 1/2 --steps 6 --warmup 3; fi
