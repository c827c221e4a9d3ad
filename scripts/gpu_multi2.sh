#!/bin/bash
# gpurun --gpus G --timeout 1500 -- bash scripts/gpu_multi2.sh G : multi-device tests, NCCL sharded-step check, torchrun bench at G ranks
set -u
G=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus_$G.txt
nvidia-smi topo -m 2>&1 | head -12 | tee -a $OUT/gpus_$G.txt
echo "== pytest two-device tests"; timeout 600 python -m pytest tests -q -m gpu -k "two_device" 2>&1 | tail -5 | tee $OUT/pytest_two_device.txt
echo "== NCCL sharded convection vs single GPU"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 \
   scripts/dist_step_check.py 100000 2>&1 | tail -2 | tee $OUT/dist_step_check_g$G.json
for N in ${SIZES:-1048576}; do
  echo "== bench $G GPUs N=$N"
  EXTRA=""; if [ "$N" -ge 16000000 ]; then EXTRA="--steps 1 --warmup 1 --e2e-steps 1 --e2e-no-warmup"; elif [ "$N" -ge 4000000 ]; then EXTRA="--steps 2 --warmup 1 --e2e-steps 1"; fi
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $G --particles $N --no-cpu $EXTRA 2>&1 | tail -1 | tee $OUT/bench_g${G}_n$N.json
done
