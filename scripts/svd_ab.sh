#!/bin/bash
# A/B of the Jacobi SVD launch variants on the C3 MPS workload (run on the GPU box).
mkdir -p gpurun_out
: > gpurun_out/svd_ab.jsonl
for cfg in ${SVD_AB_CONFIGS:-"svd_reg=4" "svd_reg=5"}; do
  args=""
  for kv in $cfg; do args="$args --opt $kv"; done
  timeout 300 python scripts/mps_bench.py --phases $args >> gpurun_out/svd_ab.jsonl 2>> gpurun_out/svd_ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/svd_ab.jsonl"):
    d = json.loads(l)
    print(d["options"], round(d["gates_per_s"], 1), d["stats"]["svd_sweeps"], round(d["stats"]["svd_seconds"], 3), round(d["stats"].get("lq_seconds", 0), 3), d["norm"])
PY
for r in 4 5; do python scripts/svd_bench.py --opt svd_reg=$r; done | tee gpurun_out/svd_bench_ab.jsonl
