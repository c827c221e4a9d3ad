#!/bin/bash
# round 2, second GPU call: the full GPU suite with the new host path / status / body / example tests, candidate cubins at 256 K.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -45 > gpurun_out/r2b_pytest.txt
K=omega3d_b200/csrc/microbench
for rep in 1 2; do
KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=$(ls kb_variants/*.cubin | tr '\n' ':') timeout 300 $K/kbench 262144 5 >> gpurun_out/r2b_kbench_cubins_256k.txt 2>&1
done
timeout 600 python bench.py --n 1048576 --steps 3 --warmup 3 > gpurun_out/r2b_bench_1m.json 2> gpurun_out/r2b_bench_1m.err
cat gpurun_out/r2b_pytest.txt; cat gpurun_out/r2b_kbench_cubins_256k.txt; head -c 400 gpurun_out/r2b_bench_1m.json; tail -3 gpurun_out/r2b_bench_1m.err
