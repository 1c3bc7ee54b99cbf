#!/bin/bash
# Development visit: per-op tests of the tensor-core kernels, then the bench with and without the halo schedule.
set -u
mkdir -p gpurun_out
echo "== tc op tests"; timeout 600 python -m pytest tests/test_gpu_tc_ops.py -q ${PYTEST_ARGS:-} > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_tc.log
for H in ${HALO_LIST:-1 0}; do
  echo "== bench HALO=$H"
  UCDIR_TC_HALO=$H timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu --dump-ops gpurun_out/ops_halo$H.json > gpurun_out/bench_halo$H.json 2> gpurun_out/bench_halo$H.err; echo "rc=$?"; tail -2 gpurun_out/bench_halo$H.err; cut -c1-400 gpurun_out/bench_halo$H.json
  python scripts/ops_summary.py gpurun_out/ops_halo$H.json 2>/dev/null | head -16
done
if [ "${E2E:-1}" = "1" ]; then
echo "== bf16 end-to-end tests"; timeout 900 python -m pytest tests/test_gpu_bf16.py -q > gpurun_out/pytest_bf16.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_bf16.log
fi
