# quick GPU pass: parity tests + bench (no profiler)
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 200 --warmup 10 ${BENCH_FLAGS:---no-cpu-baseline} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -5 gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('VALUE', d['value']/1e9, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9); r=d.get('registrations'); print('REG', r and {k:v for k,v in r.items() if k not in ('what',)}); print('PRE', d.get('preprocess')); print('STAGES', d.get('stages'))"
