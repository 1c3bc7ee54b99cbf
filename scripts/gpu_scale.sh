#!/bin/bash
# Multi-GPU visit: NCCL sharding check + scaling bench at N = 1, 2, ... NGPU.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -12 gpurun_out/topo.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
    else
      NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_$N.%h.%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
    fi
    echo "N=$N rc=$?"; tail -2 gpurun_out/scale_$N.err; cut -c1-260 gpurun_out/scale_$N.json
  fi
done
