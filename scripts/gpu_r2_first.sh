#!/bin/bash
# round 2, first GPU call: correctness of the stream-K kernels (full GPU suite), candidate cubins and staging variants at
# 256 K (kbench), short bench lines at 1 M and 4 M.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
K=omega3d_b200/csrc/microbench
for rep in 1 2; do
KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=$(ls kb_variants/*.cubin | tr '\n' ':') timeout 300 $K/kbench 262144 5 >> gpurun_out/r2a_kbench_cubins_256k.txt 2>&1
done
KBENCH_PRODUCT_ONLY=1 timeout 200 $K/kbench 262144 5 > gpurun_out/r2a_kbench_stage_256k.txt 2>&1
KBENCH_PRODUCT_ONLY=1 timeout 200 $K/kbench_cpasync 262144 5 >> gpurun_out/r2a_kbench_stage_256k.txt 2>&1
KBENCH_PRODUCT_ONLY=1 timeout 200 $K/kbench_ldgsts 262144 5 >> gpurun_out/r2a_kbench_stage_256k.txt 2>&1
timeout 600 python bench.py --n 1048576 --steps 5 --warmup 3 > gpurun_out/r2a_bench_1m.json 2> gpurun_out/r2a_bench_1m.err
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_bench_4m.json 2> gpurun_out/r2a_bench_4m.err
tail -5 gpurun_out/r2a_pytest.txt
cat gpurun_out/r2a_kbench_cubins_256k.txt gpurun_out/r2a_kbench_stage_256k.txt
head -c 600 gpurun_out/r2a_bench_1m.json; echo; head -c 600 gpurun_out/r2a_bench_4m.json
