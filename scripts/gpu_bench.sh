#!/bin/bash
# Bench only (+ per-op profile, + ncu launch list of our kernels).  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --dump-ops gpurun_out/ops_profile.json ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -3 gpurun_out/bench.err ; cat gpurun_out/bench.json
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(tc_conv|mix_halo|dense_halo|final_halo|conv_f32|sgemm_f32|softmax_rows|guidance|time_embed|gather_tiles|scatter|crop_tiles|maxpool2|gn_|cast_|layout)" -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
fi
