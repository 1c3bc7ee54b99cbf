#!/bin/bash
# ncu --set full capture (with SASS-level stall sampling) of the first dense-halo and mix-halo launches of a step.
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-(mix_halo|dense_halo)_kernel}" -s ${SKIP:-0} -c ${COUNT:-2} -o gpurun_out/prof_halo -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_halo.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_halo.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
