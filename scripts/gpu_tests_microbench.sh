# fp64 / XU pipe microbenchmark (profiles/r01_fp64_xu_pipe_microbench.txt)
mkdir -p gpurun_out
cd scripts/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_pipe fp64_pipe.cu && /tmp/fp64_pipe | tee ../../gpurun_out/fp64_pipe.txt
