#!/bin/bash
# MLP-speculator training launch (reference scripts/train_speculator.sh).
set -euo pipefail
cd "$(dirname "$0")/.."
MODEL_ARGS="\
--model_path=${MODEL_PATH:-/path/to/hf/llama} \
--model_arch=embedllama \
--model_variant=${MODEL_VARIANT:-7b} \
--ckpt_load_path=${CKPT:-/fsx/output/spec_ckpt} \
--ckpt_save_path=${CKPT:-/fsx/output/spec_ckpt} \
--sharding_strategy=${SHARDING:-tp} \
--tp_size=${TP_SIZE:-8} \
--data_path=${DATA_PATH:-/fsx/data} \
--seq_length=4096 --batch_size=2 \
--n_speculator_heads=3 --speculator_width=4096 \
--stage2_start_step=15000 --stage2_batch_size=96 --stage2_prompt_length=64 --stage2_seq_length=256 \
--num_steps=21000 --learning_rate=1e-3 --report_interval=100 --checkpoint_interval=5000"
torchrun --nnodes=1 --nproc_per_node="${GPUS_PER_NODE:-8}" --master_addr=127.0.0.1 --master_port="${MASTER_PORT:-29501}" \
    speculator/train_speculator.py ${MODEL_ARGS} "$@"
