# compute-sanitizer passes over the parity tests (memcheck on everything K3/K4/K5/K6; racecheck + synccheck on the K3 tests)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_k3_gpu.py tests/test_k4_neighbours_gpu.py tests/test_register_gpu.py tests/test_filter_gpu.py tests/test_k1_k2_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/sanitize_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_k3_gpu.py -m gpu -x -q -k "fused_per_segment or emit_matches or degenerate" 2>&1 | tail -15 | tee gpurun_out/sanitize_racecheck.txt
