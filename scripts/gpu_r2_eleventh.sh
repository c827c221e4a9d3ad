#!/bin/bash
# round 2, eleventh GPU call: the kernels with the measured statement orders: GPU suite, product cubin vs the round-1 kernel,
# the three staging variants (TMA bulk / cp.async / LDG->STS) on the final structure, bench at 1 M and 4 M.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2l_pytest.txt
cat gpurun_out/r2l_pytest.txt
O=gpurun_out/r2l_kbench.txt; : > $O
K=omega3d_b200/csrc/microbench
for n in 262144 1048576; do
  echo "== round-1 kernel (own harness) N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/r1/tuned.cubin timeout 300 kb_variants/r1/kbench $n 3 2>&1 | grep "cubin" >> $O
  echo "== product N=$n" >> $O
  KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/product.cubin timeout 300 $K/kbench $n 3 2>&1 | grep "cubin" >> $O
  KBENCH_NO_UNIFORM=1 KBENCH_CUBIN_ONLY=1 KBENCH_CUBIN=kb_variants/product.cubin timeout 300 $K/kbench $n 3 2>&1 | grep "cubin" | sed 's/cubin /per-particle-radius path: cubin /' >> $O
done
echo "== staging variants, linked (unpatched) kernels, N=262144: stage0 = cp.async.bulk (TMA), stage1 = cp.async 16 B, stage2 = LDG -> STS" >> $O
for b in kbench kbench_cpasync kbench_ldgsts; do KBENCH_PRODUCT_ONLY=1 timeout 200 $K/$b 262144 3 2>&1 | grep packed >> $O; done
cat $O
timeout 600 python bench.py --particles 1048576 --steps 3 --warmup 3 > gpurun_out/r2l_bench_1m.json 2> gpurun_out/r2l_bench_1m.err
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2l_bench_4m.json 2> gpurun_out/r2l_bench_4m.err
python - <<'PY'
import json
for f in ("gpurun_out/r2l_bench_1m.json","gpurun_out/r2l_bench_4m.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["ok"], d["cpu_baseline"]["value"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 gpurun_out/r2l_bench_4m.err
