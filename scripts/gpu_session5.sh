#!/bin/bash
# Session-5 evidence in one gpurun call: all GPU tests, smoke, both bench arms at the default size, the 4M north-star
# point with the tuned kernel, the ncu launch list of the bench command and one full ncu capture of an alternate-core
# kernel (Rosenhead-Moore, vel+grad, 256K). Usage: gpurun --timeout 900 -- bash scripts/gpu_session5.sh
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.json
echo "== bench 4M"; timeout 600 python bench.py --particles 4194304 --steps 1 --warmup 3 --e2e-steps 1 --e2e-no-warmup --cpu-seconds 4 2>&1 | tail -1 | tee $OUT/bench_4m.json
echo "== ncu launch list (bench command, 2 steps)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_1m.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full capture of ppc_kernel (Rosenhead-Moore, vel+grad, N = 256K)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ppc_kernel -c 1 -f -o $OUT/ppc_rm_full_256k \
    python tests/perf/bench_cores.py 262144 > $OUT/ncu_ppc.log 2>&1
ncu -i $OUT/ppc_rm_full_256k.ncu-rep --page raw --csv > $OUT/ppc_rm_full_256k_raw.csv 2>/dev/null
ls -la $OUT | tail -14
