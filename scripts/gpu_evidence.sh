#!/bin/bash
# evidence for the shipped particles -> points kernel (one persistent 384-thread CTA per SM, SASS-patched cubin):
#   ncu --set full at 1 M, DRAM bytes of one launch of the device-resident step at the bench shapes (-> profiles/pp2_dram_traffic.json,
#   profiles/r02_pp2_dram_traffic.txt), the launch list of the default bench command. gpurun --timeout 2400 -- bash scripts/gpu_evidence.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pp2_kernel -c 1 -f -o gpurun_out/r2u_pp2_1m \
    python tests/perf/one_call.py 1048576 > gpurun_out/r2u_ncu_1m.log 2>&1
ncu -i gpurun_out/r2u_pp2_1m.ncu-rep --page raw --csv > gpurun_out/r2u_pp2_1m_raw.csv 2>/dev/null
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum
for shape in "1048576 1048576" "4194304 4194304" "4194304 2097152" "4194304 1048576" "4194304 524288" "1048576 131072" "16777216 2097152"; do
  set -- $shape
  timeout 900 ncu --metrics $M --clock-control none -k regex:pp2_kernel -c 1 --csv --log-file gpurun_out/r2u_dram_$1x$2.csv \
      python tests/perf/one_call.py $1 $2 dev > gpurun_out/r2u_dram_$1x$2.log 2>&1
done
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches_bench_4m.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/r2u_bench_under_ncu.log 2>&1
ls -la gpurun_out/r2u_*; tail -2 gpurun_out/r2u_dram_*.csv
