#!/bin/bash
# round 2, seventh GPU call: one persistent 384-thread CTA per SM (SASS-patched), the whole GPU suite, the round-1 kernel on the
# same box for reference, bench.py at 1 M and 4 M.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2g_pytest.txt
cat gpurun_out/r2g_pytest.txt
O=gpurun_out/r2g_kbench.txt; : > $O
K=omega3d_b200/csrc/microbench
for n in 262144 1048576; do
  echo "== r1 kernel N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/r1/tuned.cubin timeout 300 kb_variants/r1/kbench $n 3 2>&1 | grep "cubin" >> $O
  echo "== 384-thread CTA N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/b384.cubin timeout 300 $K/kbench $n 3 2>&1 | grep "cubin" >> $O
done
cat $O
timeout 600 python bench.py --n 1048576 --steps 3 --warmup 3 > gpurun_out/r2g_bench_1m.json 2> gpurun_out/r2g_bench_1m.err
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2g_bench_4m.json 2> gpurun_out/r2g_bench_4m.err
python - <<'PY'
import json
for f in ("gpurun_out/r2g_bench_1m.json","gpurun_out/r2g_bench_4m.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["ok"], d["cpu_baseline"]["value"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 gpurun_out/r2g_bench_4m.err
