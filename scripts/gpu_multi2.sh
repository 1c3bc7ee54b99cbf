set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 300 --timeout-method thread > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/pytest_multi.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err; echo "N=2 rc=$?"; tail -2 gpurun_out/scale_2.err | cut -c1-300; cut -c1-400 gpurun_out/scale_2.json
