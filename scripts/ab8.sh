#!/bin/bash
# A/B of engine knobs on N GPUs (one call): usage scripts/ab8.sh N "NAME1:ENV=V ENV2=V" "NAME2:..." ...
N=$1; shift
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  env $envs timeout 300 $T bench.py --gpus $N --steps ${STEPS:-6} --warmup 3 > gpurun_out/ab_${N}_${name}.log 2>&1
  echo "== $name [$envs] rc=$?"
  grep -o '{"metric.*' gpurun_out/ab_${N}_${name}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['loss'], d['grad_norm'])"
done
