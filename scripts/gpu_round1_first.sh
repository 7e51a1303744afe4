set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 200 --warmup 10 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
