#!/bin/bash
# The PUBLIC entry points on the GPU (the calls a user of the reference makes): Llama pre-training with auto-resume,
# the Mamba entry point, the HF exporter, and -- with >= 2 GPUs -- the two-stage speculator trainer on a TP=2 HF Llama.
#   scripts/gpu_entrypoints.sh [N_GPUS]
set -u
cd "$(dirname "$0")/.."
N=${1:-1}
O=gpurun_out; mkdir -p $O
CK=/tmp/ck_llama; rm -rf $CK /tmp/ck_mamba /tmp/ck_spec /tmp/hf_tiny /tmp/hf_export
COMMON="--use_dummy_dataset=True --report_interval=4 --sharding_strategy=fsdp --use_torch_compile=False --seq_length=4096 --batch_size=2"
python main_training_llama.py --model_variant=llama2_1.4b $COMMON --low_cpu_fsdp=True --num_steps=8 --checkpoint_interval=8 \
    --ckpt_save_path=$CK --ckpt_load_path=$CK > $O/entry_llama.log 2>&1; echo "llama rc=$?"
# restart: auto-resume from the save directory (model + optimizer + step), 4 more steps
python main_training_llama.py --model_variant=llama2_1.4b $COMMON --low_cpu_fsdp=True --num_steps=12 --checkpoint_interval=100 \
    --ckpt_save_path=$CK --ckpt_load_path=$CK > $O/entry_llama_resume.log 2>&1; echo "llama resume rc=$?"
python fms_to_hf_llama.py --model_variant=llama2_1.4b --nocompiled --load_path=$CK/checkpoints/step_8_ckp --save_path=/tmp/hf_export \
    > $O/entry_export.log 2>&1; echo "export rc=$?"; ls /tmp/hf_export >> $O/entry_export.log 2>&1
python main_training_mamba.py --model_variant=mamba_2.8b $COMMON --num_steps=4 --report_interval=2 --checkpoint_interval=100 \
    --ckpt_save_path=/tmp/ck_mamba --ckpt_load_path=/tmp/ck_mamba > $O/entry_mamba.log 2>&1; echo "mamba rc=$?"
grep -hE "^(step|loss|current token per gpu per sec|device step time|model TFLOP)" $O/entry_llama.log $O/entry_llama_resume.log $O/entry_mamba.log | head -60
if [ "$N" -ge 2 ]; then
  python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import fms_to_hf_llama as ex
from fms_fsdp_b200.models.llama import LLaMA, LLaMAConfig
torch.manual_seed(0)
m = LLaMA(LLaMAConfig(src_vocab_size=32000, emb_dim=1024, nheads=8, kvheads=4, nlayers=4, multiple_of=256, max_expected_seq_len=2048))
m.reset_parameters()
ex.convert_to_hf(m, "llama2_x").save_pretrained("/tmp/hf_tiny")
print("wrote /tmp/hf_tiny")
PY
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      speculator/train_speculator.py --model_arch=embedllama --model_variant=7b --model_path=/tmp/hf_tiny --sharding_strategy=tp \
      --tp_size=2 --use_dummy_dataset=True --seq_length=256 --batch_size=2 --num_steps=8 --stage2_start_step=4 --report_interval=2 \
      --checkpoint_interval=100 --stage2_batch_size=8 --stage2_prompt_length=16 --stage2_seq_length=32 --speculator_width=1024 \
      --n_speculator_heads=3 --ckpt_save_path=/tmp/ck_spec --ckpt_load_path=/tmp/ck_spec --use_torch_compile=False \
      > $O/entry_speculator_tp2.log 2>&1; echo "speculator rc=$?"
  grep -E "^(step|loss|loss [0-9])" $O/entry_speculator_tp2.log | head -40; tail -5 $O/entry_speculator_tp2.log
fi
