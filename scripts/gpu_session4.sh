#!/bin/bash
# One gpurun call: all GPU tests, smoke, both bench arms at the default size, the ncu launch list of the default bench
# command, one full ncu capture of the dominant kernel at the default size, and the kernel-variant shoot-out.
# Usage: gpurun --timeout 1700 -- bash scripts/gpu_session4.sh
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.json
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== kbench"; timeout 300 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | grep packed | tee $OUT/kbench.txt
echo "== kbench general-radius path"; KBENCH_NO_UNIFORM=1 timeout 300 omega3d_b200/csrc/microbench/kbench 262144 3 2>&1 | grep packed | tee $OUT/kbench_nouniform.txt
echo "== ncu launch list (default bench command, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_1m.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full capture of the dominant kernel at N = 1M"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pp2_kernel -s 1 -c 1 -f -o $OUT/pp2_full_1m \
    python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/pp2_full_1m.ncu-rep --page raw --csv > $OUT/pp2_full_1m_raw.csv 2>/dev/null
ls -la $OUT
