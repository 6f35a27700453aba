#!/bin/bash
# Single/multi-node launch (capability parity with reference scripts/train.sh; fixes its stale entry point).
# One process per GPU; rendezvous only uses c10d -- the hot collectives are the engine's NVLink peer kernels.
set -euo pipefail
NNODES=${NNODES:-1}
NODE_RANK=${NODE_RANK:-0}
MASTER_ADDR=${MASTER_ADDR:-127.0.0.1}
MASTER_PORT=${MASTER_PORT:-29500}
GPUS_PER_NODE=${GPUS_PER_NODE:-8}

MODEL_ARGS="\
--model_variant=${MODEL_VARIANT:-llama2_7b} \
--use_dummy_dataset=${USE_DUMMY:-False} \
--ckpt_load_path=${CKPT_LOAD:-/fsx/output/ckpt} \
--ckpt_save_path=${CKPT_SAVE:-/fsx/output/ckpt} \
--data_path=${DATA_PATH:-/fsx/data} \
--sharding_strategy=${SHARDING:-hsdp} \
--fsdp_activation_checkpointing=False \
--selective_checkpointing=1 \
--mixed_precision=True \
--low_cpu_fsdp=True \
--batch_size=2 \
--learning_rate=3e-4 \
--checkpoint_interval=5000 \
--tracker=${TRACKER:-None} \
--tracker_dir=${TRACKER_DIR:-/fsx/aim_logs/llama} \
--tracker_project_name=llama \
--tracker_run_id=None \
--report_interval=100"

cd "$(dirname "$0")/.."
python -c 'import __graft_entry__ as g; g.build()'
torchrun --nnodes="${NNODES}" --node_rank="${NODE_RANK}" --nproc_per_node="${GPUS_PER_NODE}" \
    --master_addr="${MASTER_ADDR}" --master_port="${MASTER_PORT}" \
    "${ENTRY:-main_training_llama.py}" ${MODEL_ARGS} "$@"
