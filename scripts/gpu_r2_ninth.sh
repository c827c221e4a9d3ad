#!/bin/bash
# round 2, ninth GPU call: whole-blocks-first + stream-K-tail partition, velocity+gradient kernel without its per-tile barrier:
# the GPU suite, kbench at 256 K / 1 M, DRAM bytes of one launch of the device-resident step at the bench shapes, bench at 4 M.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2j_pytest.txt
cat gpurun_out/r2j_pytest.txt
O=gpurun_out/r2j_kbench.txt; : > $O
K=omega3d_b200/csrc/microbench
for n in 262144 1048576; do
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=omega3d_b200/lib/pp2_tuned.cubin timeout 300 $K/kbench $n 3 2>&1 | grep "cubin" >> $O
done
cat $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum
for shape in "1048576 1048576" "4194304 4194304" "4194304 524288" "1048576 131072" "16777216 2097152"; do
  set -- $shape
  timeout 900 ncu --metrics $M --clock-control none -k regex:pp2_kernel -c 1 --csv --log-file gpurun_out/r2j_dram_$1x$2.csv \
      python tests/perf/one_call.py $1 $2 dev > gpurun_out/r2j_dram_$1x$2.log 2>&1
  tail -n 1 gpurun_out/r2j_dram_$1x$2.log
  grep -h "dram__bytes\|gpu__time" gpurun_out/r2j_dram_$1x$2.csv | awk -F'","' '{print $(NF-2), $NF}'
done
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2j_bench_4m.json 2> gpurun_out/r2j_bench_4m.err
python - <<'PY'
import json
for f in ("gpurun_out/r2j_bench_4m.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["ok"], d["cpu_baseline"]["value"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 gpurun_out/r2j_bench_4m.err
