#!/bin/bash
# Round-end evidence in one gpurun call: all GPU tests, smoke, both bench arms at the default size, a 4M bench point,
# the ncu launch list of the default bench command and one full ncu capture of the dominant kernel (256K: 40 replays of
# a 75 ms launch instead of a 1.2 s one). Usage: gpurun --timeout 1700 -- bash scripts/gpu_final.sh
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench 4M"; timeout 900 python bench.py --particles 4194304 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_4m.json
echo "== ncu launch list (default bench command, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_1m.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full capture of the dominant kernel (N = 256K)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pp2_kernel -s 1 -c 1 -f -o $OUT/pp2_full_256k \
    python bench.py --steps 1 --warmup 1 --particles 262144 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/pp2_full_256k.ncu-rep --page raw --csv > $OUT/pp2_full_256k_raw.csv 2>/dev/null
ls -la $OUT | tail -12
