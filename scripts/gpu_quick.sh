# quick GPU pass: parity tests + bench (no profiler)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -5 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json
