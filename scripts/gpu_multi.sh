#!/bin/bash
# gpurun --gpus G --timeout 480 -- bash scripts/gpu_multi.sh G
# multi-device evidence on ONE box with G GPUs: the in-process multi-device tests (one context over several GPUs), the
# NCCL sharded-step bit-identity check, bench.py under torchrun at G ranks (per-rank parity; at G = 8 the 16 M extra step),
# and bench.py --inproc G (one process, pageable host arrays, sources replicated over NVLink).
set -u
G=${1:-2}
STEPS=${STEPS:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/r2_gpus_$G.txt
nvidia-smi topo -m 2>&1 | head -12 >> $OUT/r2_gpus_$G.txt
echo "== bench $G GPUs (torchrun, default size)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $G --steps $STEPS --warmup 3 2>&1 | tail -1 | tee $OUT/r2_bench_g${G}_4m.json
echo "== pytest multi-device tests"; timeout 200 python -m pytest tests -q -m gpu -k "two_device" 2>&1 | tail -3 | tee $OUT/r2_pytest_two_device_g$G.txt
echo "== NCCL sharded convection vs single GPU"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29533 \
   scripts/dist_step_check.py 100000 2>&1 | tail -1 | tee $OUT/r2_dist_step_check_g$G.json
echo "== bench --inproc $G (one process, pageable host arrays)"
timeout 200 python bench.py --inproc $G --steps 3 --warmup 1 --particles 1048576 2>&1 | tail -1 | tee $OUT/r2_bench_inproc${G}_1m.json
timeout 200 python bench.py --inproc $G --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/r2_bench_inproc${G}_4m.json
if [ "${ALSO_1M:-0}" = 1 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $G --steps 5 --warmup 3 --particles 1048576 --extra-n 0 2>&1 | tail -1 | tee $OUT/r2_bench_g${G}_1m.json
fi
