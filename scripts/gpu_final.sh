#!/bin/bash
# Round-end visit (1 GPU): smoke, every GPU test, the bench arms (ours / reference CPU / eager torch GPU), the other workloads
# (C2, C5, the reference-default tiling `sr.py` runs, the 16-tile rank share), the fp32-tolerance mode, the streamed-only
# ablation, an ncu launch list and ncu --set full captures of the top kernels.  Everything lands in gpurun_out/final/;
# `python scripts/collect_profiles.py r02` copies the summaries into profiles/.
set -u
O=gpurun_out/final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-300
cp gpurun_out/bf16_errors.json gpurun_out/fp32_errors.json $O/ 2>/dev/null
echo "== bench ours (default line)"; timeout 900 python bench.py --steps 20 --warmup 3 --dump-ops $O/ops_profile.json > $O/bench.json 2> $O/bench.err; echo "rc=$?"; tail -2 $O/bench.err; cut -c1-260 $O/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"; cut -c1-260 $O/bench_reference.json
echo "== bench eager torch on this GPU"; timeout 600 python bench.py --impl torch_gpu > $O/bench_torch_gpu.json 2> $O/bench_torch_gpu.err; echo "rc=$?"; cut -c1-200 $O/bench_torch_gpu.json
echo "== fp32_tc"; timeout 600 python bench.py --precision fp32_tc --steps 10 --warmup 3 --no-cpu --no-eager --dump-ops $O/ops_profile_fp32tc.json > $O/bench_fp32tc.json 2> $O/bench_fp32tc.err; echo "rc=$?"; cut -c1-200 $O/bench_fp32tc.json
for W in c2_256_b8 c5_sid_512_b32 c3_1152_ref_tiling rank_share_16_tiles; do
  echo "== $W"; timeout 600 python bench.py --workload $W --steps ${STEPS_OTHER:-5} --warmup 3 --no-cpu --dump-ops $O/ops_profile_$W.json > $O/bench_$W.json 2> $O/bench_$W.err; echo "rc=$?"; cut -c1-200 $O/bench_$W.json
done
echo "== streamed only (UCDIR_TC_HALO=0)"; UCDIR_TC_HALO=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --no-parity-mode > $O/bench_halo0.json 2> $O/bench_halo0.err; echo "rc=$?"; cut -c1-160 $O/bench_halo0.json
echo "== materialised attention (UCDIR_TC_FLASH=0) on the reference tiling"; UCDIR_TC_FLASH=0 UCDIR_CHUNK_PIXELS=2359296 timeout 400 python bench.py --workload c3_1152_ref_tiling --steps 5 --warmup 3 --no-cpu --no-eager --no-parity-mode > $O/bench_ref_tiling_flash0.json 2> $O/bench_flash0.err; echo "rc=$?"; cut -c1-160 $O/bench_ref_tiling_flash0.json
KREG="regex:^(tc_conv|mix_halo|dense_halo|final_halo|flash_attn|conv_f32|sgemm_f32|softmax_rows|guidance|time_embed|gather_tiles|scatter|crop_tiles|maxpool2|gn_|cast_|split_rows|layout|to_image)"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 4000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-parity-mode > $O/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: mix halo + dense halo + streamed conv"
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:(mix_halo|dense_halo)_kernel" -s 1 -c 2 -o $O/prof_halo -f python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-parity-mode > $O/ncu_halo.log 2>&1; echo "rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_conv_kernel<\(int\)64, \(int\)64, \(int\)256, \(int\)1, \(int\)[01]," -s 10 -c 3 -o $O/prof_tc -f python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-parity-mode > $O/ncu_tc.log 2>&1; echo "rc=$?"
echo "== ncu full: flash attention at 16384 tokens (reference tiling)"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:flash_attn_kernel" -c 1 -o $O/prof_attn -f python bench.py --workload c3_1152_ref_tiling --steps 1 --warmup 3 --no-cpu --no-eager --no-parity-mode > $O/ncu_attn.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep
