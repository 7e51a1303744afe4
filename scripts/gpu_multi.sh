# multi-GPU bench (one process per GPU, NCCL), as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -3; nvidia-smi topo -m | head -12; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_gpus$N.json 2> gpurun_out/bench_gpus$N.err; tail -5 gpurun_out/bench_gpus$N.err
python - <<PY
import json
l=json.load(open("gpurun_out/bench_gpus$N.json"))
print("N=$N value %.2f G pairs/s  ms %.4f  e2e %.2f G (%.4f ms/step)  reg %.2f M/s (%.2f ms)" % (l["value"]/1e9, l["ms_per_step"], l["e2e"]["value"]/1e9, l["e2e"]["ms_per_step"], l["registrations"]["value"]/1e6, l["registrations"]["ms_per_batch"]))
print(json.dumps(l["per_rank"]))
print(json.dumps(l["configs"]["c3"]))
PY
