#!/bin/bash
# ncu evidence for the current product kernel + the secondary panel measurements
set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== panels 320"; timeout 600 python tests/perf/bench_panels.py 2 1000000 2>&1 | tail -5 | tee $OUT/panels_320.jsonl
echo "== panels 5120"; timeout 600 python tests/perf/bench_panels.py 4 1000000 2>&1 | tail -5 | tee $OUT/panels_5120.jsonl
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --particles 262144 --no-cpu --e2e-steps 1 > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pp2_kernel -s 1 -c 1 -f -o $OUT/pp2_full \
    python bench.py --steps 1 --warmup 1 --particles 262144 --no-cpu --e2e-steps 1 > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -5
