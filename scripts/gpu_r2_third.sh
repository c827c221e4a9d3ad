#!/bin/bash
# round 2, third GPU call: the two-copy stream-K structure (correctness subset, candidate cubins, staging variants) and the
# panel work-queue kernel against the per-lane baseline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "not 4m" 2>&1 | tail -8 > gpurun_out/r2c_pytest.txt
K=omega3d_b200/csrc/microbench
for rep in 1 2; do
KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=$(ls kb_variants/*.cubin | tr '\n' ':') timeout 300 $K/kbench 262144 5 >> gpurun_out/r2c_kbench_cubins_256k.txt 2>&1
done
KBENCH_PRODUCT_ONLY=1 timeout 200 $K/kbench 262144 5 > gpurun_out/r2c_kbench_stage_256k.txt 2>&1
KBENCH_PRODUCT_ONLY=1 timeout 200 $K/kbench_cpasync 262144 5 >> gpurun_out/r2c_kbench_stage_256k.txt 2>&1
KBENCH_PRODUCT_ONLY=1 timeout 200 $K/kbench_ldgsts 262144 5 >> gpurun_out/r2c_kbench_stage_256k.txt 2>&1
for q in queue noqueue; do
  timeout 200 python tests/perf/bench_panels.py 2 1000000 $q 2>&1 | tail -7 > gpurun_out/r2c_panels_320_$q.jsonl
  timeout 200 python tests/perf/bench_panels.py 4 1000000 $q 2>&1 | tail -7 > gpurun_out/r2c_panels_5120_$q.jsonl
  timeout 200 python tests/perf/bench_panels.py 4 262144 $q 2>&1 | tail -7 > gpurun_out/r2c_panels_5120_262k_$q.jsonl
done
timeout 600 python bench.py --n 1048576 --steps 3 --warmup 3 > gpurun_out/r2c_bench_1m.json 2> gpurun_out/r2c_bench_1m.err
cat gpurun_out/r2c_pytest.txt; cat gpurun_out/r2c_kbench_cubins_256k.txt | grep velgrad; cat gpurun_out/r2c_kbench_stage_256k.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_panels_*.jsonl")):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split("/")[-1], d["routine"], "%.3f ms"%d["kernel_ms"], "err %.2e"%(d.get("max_rel_err_vs_oracle_sample") or 0))
PY
head -c 300 gpurun_out/r2c_bench_1m.json; tail -3 gpurun_out/r2c_bench_1m.err
