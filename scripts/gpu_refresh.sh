#!/bin/bash
# Short round-end refresh: smoke, every GPU test, bench (C3 + C2), ncu launch list.
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 3 --dump-ops gpurun_out/ops_profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
timeout 300 python bench.py --workload c2_256_b8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?"; cut -c1-160 gpurun_out/bench_c2.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(tc_conv|mix_halo|dense_halo|final_halo|conv_f32|sgemm_f32|softmax_rows|guidance|time_embed|gather_tiles|scatter|crop_tiles|maxpool2|gn_|cast_|layout|to_image)" -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
