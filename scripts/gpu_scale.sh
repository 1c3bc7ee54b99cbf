#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL sharding check (tile + batch sharding == unsharded) and the scaling benches:
# C3 (tile sharding, one all-gather per step) at N = 1, 2, 4, 8 and the batch workloads C5 / C2 (batch sharding, no per-step
# collective) at the box size.  Outputs under gpurun_out/scale/.
set -u
O=gpurun_out/scale
mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
nvidia-smi topo -m > $O/topo.txt 2>&1; head -12 $O/topo.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -3 $O/pytest_multi.log
run() {  # N, workload, steps, output tag
  if [ $1 -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps $3 --warmup 3 --no-cpu --no-eager --workload $2 > $O/$4_$1.json 2> $O/$4_$1.err
  else
    NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$O/nccl_$4_$1.%h.%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$((30 + $1)) bench.py --gpus $1 --steps $3 --warmup 3 --no-cpu --no-eager --workload $2 > $O/$4_$1.json 2> $O/$4_$1.err
  fi
  echo "$4 N=$1 rc=$?"; tail -1 $O/$4_$1.err | cut -c1-200; python -c "import json; d=json.load(open('$O/$4_$1.json')); print('   value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'frac_of_step', d['roofline']['frac_of_step'], d['collective'])"
}
# SKIP1=1: leave the single-GPU runs to a 1-GPU visit (an 8-GPU box is charged 8x while one GPU works)
for N in 1 2 4 8; do
  if [ $N -le $NG ] && ! { [ $N -eq 1 ] && [ "${SKIP1:-0}" = "1" ]; }; then run $N c3_1024_tile128 ${STEPS:-20} scale; fi
done
[ "${SKIP1:-0}" = "1" ] || run 1 c5_sid_512_b32 5 c5
run $NG c5_sid_512_b32 5 c5
[ "${SKIP1:-0}" = "1" ] || run 1 c2_256_b8 10 c2
run $NG c2_256_b8 10 c2
grep -h "NVLS\|Connected all" $O/nccl_scale_$NG.*.log 2>/dev/null | head -3
rm -f $O/nccl_*.log
