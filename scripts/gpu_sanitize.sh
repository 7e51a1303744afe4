# compute-sanitizer sweep over every -m gpu test: memcheck, racecheck (shared-memory hazards: the mbarrier / TMA ring, the stage-buffer
# reuse as reduction scratch, the ticket folds, K7's resident buffers, K1's sort), synccheck.  The full-length replay and the
# full-size batch are deselected from racecheck / synccheck only for run time (their kernels are the same ones the other tests launch).
mkdir -p gpurun_out
SKIP='not full_length_replay and not every_step_matches'
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 2400 compute-sanitizer --tool $tool $extra --error-exitcode 9 python -m pytest tests -m gpu -q -k "$SKIP" -p no:cacheprovider 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_sanitizer_$tool.txt
  echo "== $tool: exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
