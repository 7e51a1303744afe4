# multi-GPU bench (one process per GPU, NCCL), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_gpus$N.json 2> gpurun_out/bench_gpus$N.err; tail -5 gpurun_out/bench_gpus$N.err; cat gpurun_out/bench_gpus$N.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>&1 | tail -2 | cut -c1-600
