# rebuild K3 with different pipeline depths / occupancy targets on the GPU box and bench each (tuning aid)
mkdir -p gpurun_out
: > gpurun_out/k3_sweep.txt
for cfg in ${K3_CFGS:-"2 4 128" "3 5 96"}; do
  set -- $cfg
  RANDT_NVCC_FLAGS="-DRANDT_K3_STAGES=$1 -DRANDT_K3_MIN_CTAS=$2 -DRANDT_K3_THREADS=${3:-128}" python -m randt_slam_b200.build --force > gpurun_out/k3_sweep_build.log 2>&1 || tail -5 gpurun_out/k3_sweep_build.log
  v=$(timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --reg-steps 0 --pre-scans 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'])")
  echo "stages=$1 min_ctas=$2 threads=${3:-128} : Gpairs/s ms frac = $v" | tee -a gpurun_out/k3_sweep.txt
done
