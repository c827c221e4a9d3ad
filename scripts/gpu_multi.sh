#!/bin/bash
# gpurun --gpus G --timeout 1500 -- bash scripts/gpu_multi.sh G [n] : GPU tests + the torchrun bench at G ranks
set -u
G=${1:-2}; N=${2:-1048576}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus_$G.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu_multi.txt
echo "== bench 1 GPU"; timeout 600 python bench.py --particles $N --no-cpu 2>&1 | tail -1 | tee $OUT/bench_g1_n$N.json
echo "== bench $G GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $G --particles $N --no-cpu 2>&1 | tail -3 | tee $OUT/bench_g${G}_n$N.json
