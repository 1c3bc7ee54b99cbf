set -u
mkdir -p gpurun_out
UCDIR_PRECISION=bf16 timeout 200 python scripts/e2e_breakdown.py c3_1024_tile128 2>&1 | grep -v Warning | tail -14
timeout 400 python bench.py --steps 20 --warmup 3 --dump-ops gpurun_out/ops_profile.json > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench.json
timeout 300 python bench.py --workload c2_256_b8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_c2.json
python -c "
import json
for f in ['gpurun_out/bench.json','gpurun_out/bench_c2.json']:
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
