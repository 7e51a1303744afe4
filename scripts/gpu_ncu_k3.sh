# ncu full capture of the K3 fused kernel + launch list (bench workload)
set -x
mkdir -p gpurun_out
TAG=${1:-k3}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_fused -s 3 -c 2 -o gpurun_out/${TAG} -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_full.log 2>&1; tail -2 gpurun_out/${TAG}_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch.log 2>&1; tail -1 gpurun_out/${TAG}_launch.log
