#!/bin/bash
# Development visit (tightly time-boxed): per-op tests of the halo kernels, then -- only if they pass -- the bench.
set -u
mkdir -p gpurun_out
echo "== tc op tests"
timeout ${T_TESTS:-240} python -m pytest tests/test_gpu_tc_ops.py -q -x --timeout 60 --timeout-method thread ${PYTEST_ARGS:--k "halo or final"} > gpurun_out/pytest_tc.log 2>&1
rc=$?; echo "rc=$rc"; grep -v "mbarrier timeout" gpurun_out/pytest_tc.log | tail -${TAIL:-40} | cut -c1-600
grep "mbarrier timeout" gpurun_out/pytest_tc.log | sed 's/thread [0-9]*/thread N/' | sort | uniq -c | head -20
if [ $rc -ne 0 ] && [ "${FORCE_BENCH:-0}" != "1" ]; then exit 0; fi
for H in ${HALO_LIST:-1}; do
  echo "== bench HALO=$H"
  UCDIR_TC_HALO=$H timeout ${T_BENCH:-240} python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu --dump-ops gpurun_out/ops_halo$H.json > gpurun_out/bench_halo$H.json 2> gpurun_out/bench_halo$H.err; echo "rc=$?"; tail -2 gpurun_out/bench_halo$H.err | cut -c1-300; grep -v "mbarrier timeout" gpurun_out/bench_halo$H.json | cut -c1-400
  python scripts/ops_summary.py gpurun_out/ops_halo$H.json 2>/dev/null | head -24
done
if [ "${E2E:-0}" = "1" ]; then
echo "== bf16 end-to-end tests"; timeout ${T_E2E:-400} python -m pytest tests/test_gpu_bf16.py -q -x --timeout 120 --timeout-method thread > gpurun_out/pytest_bf16.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_bf16.log | cut -c1-400
fi
