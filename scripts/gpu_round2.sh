#!/bin/bash
# gpurun --timeout 2400 -- bash scripts/gpu_round2.sh    - the recipe behind the round-2 one-GPU numbers of DESIGN.md: GPU suite,
# product cubin against the round-1 kernel (kb_variants/r1: `git show cb1791d` sources built with its own kbench), the three
# staging variants (TMA bulk / cp.async / LDG->STS; make -C omega3d_b200/csrc kbench kbench_stage), bench at 1 M and 4 M.
# -> profiles/r02_pytest_gpu.txt, r02_kbench_product_vs_r1.txt, r02_bench_1gpu_{1m,4m}.json
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
