set -u
mkdir -p gpurun_out
echo "== bf16 end-to-end tests"; timeout 500 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -q -x --timeout 180 --timeout-method thread > gpurun_out/pytest_e2e.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_e2e.log | cut -c1-400
echo "== ncu"; timeout 400 ncu --set full --clock-control none --import-source on -k "regex:(mix_halo|dense_halo)_kernel" -c 2 -o gpurun_out/prof_halo -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_halo.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_halo.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
