# Final pass of round 2 on one B200 after the window solver went in (the kernels K1 / K3 / K7 / K8 are unchanged since gpu_r02_profile.sh's
# ncu captures): every GPU parity test, smoke(), bench.py (+ the reference arm).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -3 gpurun_out/r02_bench_final.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; tail -2 gpurun_out/r02_bench_ref.err
g++ -std=c++17 -O1 -Iinclude examples/local_fuser_flow.cpp -Lrandt_slam_b200 -lrandt_host -lrandt_gpu -Wl,-rpath,$PWD/randt_slam_b200 -o /tmp/local_fuser_flow && /tmp/local_fuser_flow > gpurun_out/r02_example_output.txt 2>&1; tail -12 gpurun_out/r02_example_output.txt
