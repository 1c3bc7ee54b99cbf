#!/bin/bash
# Ablation of the mix-conv epilogue (bench only; results are wrong by construction, tests are not run).
set -u
mkdir -p gpurun_out
for A in 1 2; do
  UCDIR_NVCC_EXTRA="-DUCDIR_ABLATE=$A" python -m ucdir_b200.build --force > gpurun_out/build_ablate_$A.log 2>&1
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --dump-ops gpurun_out/ops_ablate_$A.json > gpurun_out/bench_ablate_$A.json 2> gpurun_out/bench_ablate_$A.err
  echo "ablate $A rc=$?"; cut -c1-160 gpurun_out/bench_ablate_$A.json
done
