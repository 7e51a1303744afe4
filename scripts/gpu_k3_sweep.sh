# rebuild K3 with different pipeline depths / occupancy targets on the GPU box and bench each (tuning aid)
mkdir -p gpurun_out
: > gpurun_out/k3_sweep.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in ${K3_CFGS:-"2 4" "2 5" "2 6"}; do
  set -- $cfg
  RANDT_NVCC_FLAGS="-DRANDT_K3_STAGES=$1 -DRANDT_K3_MIN_CTAS=$2" python -m randt_slam_b200.build --force > /dev/null 2>&1
  v=$(timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'])")
  echo "stages=$1 min_ctas=$2 : Gpairs/s ms frac = $v" | tee -a gpurun_out/k3_sweep.txt
done
