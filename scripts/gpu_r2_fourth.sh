#!/bin/bash
# round 2, fourth GPU call: drop-in body step + panel tests, single-copy vs two-copy walk candidates (kbench), panel queue
# kernels (both directions) against the per-lane baselines, ncu --set full of pan_pts_queue_kernel<true> and reflect_kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "dropin or sphere or panel or pan_ or golden or bem or reflect" 2>&1 | tail -8 > gpurun_out/r2d_pytest.txt
K=omega3d_b200/csrc/microbench
for rep in 1 2; do
KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=$(ls kb_variants/*.cubin | tr '\n' ':') timeout 300 $K/kbench 262144 5 >> gpurun_out/r2d_kbench_cubins_256k.txt 2>&1
done
for q in queue noqueue; do
  timeout 200 python tests/perf/bench_panels.py 2 1000000 $q 2>&1 | tail -7 > gpurun_out/r2d_panels_320_$q.jsonl
  timeout 200 python tests/perf/bench_panels.py 4 1000000 $q 2>&1 | tail -7 > gpurun_out/r2d_panels_5120_$q.jsonl
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pan_pts_queue_kernel -s 2 -c 1 -f -o gpurun_out/r2d_pan_pts_queue \
    python tests/perf/bench_panels.py 4 262144 > gpurun_out/r2d_ncu_pan.log 2>&1
ncu -i gpurun_out/r2d_pan_pts_queue.ncu-rep --page raw --csv > gpurun_out/r2d_pan_pts_queue_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pts_pan_queue_kernel -s 1 -c 1 -f -o gpurun_out/r2d_pts_pan_queue \
    python tests/perf/bench_panels.py 4 262144 > gpurun_out/r2d_ncu_ptspan.log 2>&1
ncu -i gpurun_out/r2d_pts_pan_queue.ncu-rep --page raw --csv > gpurun_out/r2d_pts_pan_queue_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:reflect_kernel -s 1 -c 1 -f -o gpurun_out/r2d_reflect \
    python tests/perf/bench_panels.py 2 1000000 > gpurun_out/r2d_ncu_reflect.log 2>&1
ncu -i gpurun_out/r2d_reflect.ncu-rep --page raw --csv > gpurun_out/r2d_reflect_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/r2d_pytest.txt
grep "velgrad" gpurun_out/r2d_kbench_cubins_256k.txt | sed 's/.*cubin \(kb_variants[^ ]*\).*grid= *444 *\([0-9.]*\) ms.*/\1 \2/' | sort -k2 -n
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_panels_*.jsonl")):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: continue
        print(f.split("/")[-1], d["routine"], "%.3f ms"%d["kernel_ms"], "err %.2e"%(d.get("max_rel_err_vs_oracle_sample") or 0))
PY
ls -la gpurun_out/r2d_*raw.csv
