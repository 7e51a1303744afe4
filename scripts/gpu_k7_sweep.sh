# K7 tuning aid: rebuild with other CTA shapes (warps per CTA : CTAs per SM) on the box and time the registrations leg of the bench
set -x
mkdir -p gpurun_out
for cfg in "${@:-4:4}"; do
  IFS=: read wp ct <<< "$cfg"
  RANDT_NVCC_FLAGS="-DRANDT_K7_WARPS=$wp -DRANDT_K7_MIN_CTAS=$ct" python -c "from randt_slam_b200 import build; build.build_all(force=True)" || continue
  timeout 300 python -m pytest tests/test_register_gpu.py tests/test_literal_batch_gpu.py -m gpu -x -q 2>&1 | tail -1
  timeout 300 python bench.py --no-cpu-baseline --pre-scans 0 --no-configs --steps 20 --reg-streams 0 | python -c "
import json,sys; l=json.loads(sys.stdin.read()); r=l['registrations']; print('K7SWEEP warps=$wp ctas=$ct  %.3f M reg/s  %.3f ms per batch' % (r['value']/1e6, r['ms_per_batch']))"
done
