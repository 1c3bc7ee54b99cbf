#!/bin/bash
# compute-sanitizer over the kernel unit tests (SURVEY 5 "race detection / sanitizers"): memcheck on every per-op tensor-core test,
# racecheck (shared-memory hazards between the TMA / MMA / epilogue roles) and synccheck on the hand-rolled mbarrier protocols.
# Runs on the GPU box:  gpurun --timeout 2400 -- 'bash scripts/run_sanitizer.sh'   ->  gpurun_out/sanitizer/*.log
# The library is rebuilt with a long mbarrier timeout first: under the sanitizer a kernel runs 10-100x slower and the product's
# 2-second deadlock trap would fire.
set -u
OUT=gpurun_out/sanitizer
mkdir -p $OUT
UCDIR_NVCC_EXTRA=-DUCDIR_MBAR_TIMEOUT_NS=900000000000ull python -m ucdir_b200.build --force > $OUT/build.log 2>&1 || { echo "build failed"; tail -5 $OUT/build.log; exit 1; }
SAN=/usr/local/cuda/bin/compute-sanitizer
SMALL='dense_halo or grouped_mix or final or flash_attention or strided or upsample or attention_chain or row3 or split'
run() {  # tool, -k expression, log name, time limit
  timeout "$4" $SAN --tool "$1" --error-exitcode 86 --print-limit 20 python -m pytest tests/test_gpu_tc_ops.py -x -q -m gpu -k "$2" > $OUT/$3.log 2>&1
  echo "$1 [$2]: exit $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/$3.log | tr '\n' ' ')"
}
run memcheck "not 16k_tokens and not 2176" memcheck_tc_ops 700s
run racecheck "$SMALL and not 16k_tokens and not 2176" racecheck_tc_ops 600s
run synccheck "flash_attention and not 16k_tokens and not 2176" synccheck_attn 300s
python -m ucdir_b200.build --force > /dev/null 2>&1      # restore the product build (2 s trap)
