# K3 tuning aid: rebuild with other pipeline depths / CTAs per SM on the box and bench each (the default build runs the parity tests first)
set -x
mkdir -p gpurun_out
for cfg in "${@:-2:4:128}"; do
  IFS=: read st ct th <<< "$cfg"; th=${th:-128}
  RANDT_NVCC_FLAGS="-DRANDT_K3_STAGES=$st -DRANDT_K3_MIN_CTAS=$ct -DRANDT_K3_THREADS=$th" python -c "from randt_slam_b200 import build; build.build_all(force=True)" || continue
  timeout 300 python -m pytest tests/test_k3_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2
  timeout 300 python bench.py --no-cpu-baseline --reg-steps 0 --pre-scans 0 --no-configs > gpurun_out/sweep_${st}_${ct}_${th}.json 2> gpurun_out/sweep_${st}_${ct}_${th}.err
  python -c "
import json; d=json.load(open('gpurun_out/sweep_${st}_${ct}_${th}.json')); print('SWEEP stages=$st ctas=$ct threads=$th VALUE %.2f G pairs/s  %.2f us  frac %.3f  e2e %.2f' % (d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']/1e9))"
done
