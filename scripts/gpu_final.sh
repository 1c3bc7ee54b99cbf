#!/bin/bash
# Round-end visit: smoke, every GPU test, both bench arms, C2 workload, ncu launch list, ncu --set full captures of the halo kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench ours"; timeout 600 python bench.py --steps 20 --warmup 3 --dump-ops gpurun_out/ops_profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
echo "== bench C2"; timeout 400 python bench.py --workload c2_256_b8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_c2.json
echo "== bench streamed only (UCDIR_TC_HALO=0)"; UCDIR_TC_HALO=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_halo0.json 2> gpurun_out/bench_halo0.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_halo0.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(tc_conv|mix_halo|dense_halo|final_halo|conv_f32|sgemm_f32|softmax_rows|guidance|time_embed|gather_tiles|scatter|crop_tiles|maxpool2|gn_|cast_|layout|to_image)" -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: dense halo + mix halo"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:(mix_halo|dense_halo)_kernel" -s 1 -c 2 -o gpurun_out/prof_halo -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_halo.log 2>&1; echo "rc=$?"
echo "== ncu full: final halo"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:final_halo_kernel" -c 1 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_final.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
