# Round-1 GPU pass: parity tests, smoke, bench (+ reference arm), ncu launch list + full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r01_v3}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_${TAG}.json 2>&1; cat gpurun_out/bench_ref_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reg-steps 1 --reg-streams 0 --pre-scans 1 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_fused -s 3 -c 2 -o gpurun_out/${TAG}_k3_fused -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reg-steps 0 --pre-scans 0 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
