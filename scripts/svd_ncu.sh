#!/bin/bash
# SVD microbench without the profiler, then one ncu --set full capture of the cross-step kernel (sweep 2).
mkdir -p gpurun_out
python scripts/svd_bench.py --reps 1 --max-sweeps 4 > gpurun_out/svd_bench4.jsonl 2> gpurun_out/svd_bench.err || { tail -5 gpurun_out/svd_bench.err; exit 1; }

cat gpurun_out/svd_bench4.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_cross_dual -s 150 -c 2 \
  -o gpurun_out/prof_svd_dual -f python scripts/svd_bench.py --reps 1 --max-sweeps 4 > gpurun_out/ncu_svd.log 2>&1
tail -3 gpurun_out/ncu_svd.log
ncu -i gpurun_out/prof_svd_dual.ncu-rep --page raw --csv > gpurun_out/prof_svd_dual_raw.csv 2>/dev/null
ls -la gpurun_out/prof_svd_dual*
