# parity tests + bench (no profiler) + one full ncu capture of the K3 fused kernel
set -x
mkdir -p gpurun_out
TAG=${1:-k3}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -5 gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('VALUE', d['value']/1e9, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_fused -s 3 -c 1 -o gpurun_out/${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1; tail -1 gpurun_out/${TAG}_full.log
