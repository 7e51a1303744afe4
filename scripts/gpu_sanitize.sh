# compute-sanitizer sweep over every -m gpu test: memcheck, racecheck (shared-memory hazards: the mbarrier / TMA ring, the stage-buffer
# reuse as reduction scratch, the ticket folds, K7's resident buffers and its team-mode exchange, K1's shared-memory sort, the fused
# single-map association), synccheck.  The full-length replay, the every-step teacher-forced chain and the 600-registration mode
# comparison are deselected for run time only (their kernels are the same ones the other tests launch: test_scan_step_equals_the_separate_calls
# runs K1 / fused K2 / team K7 / merge on 40 scans, test_literal_256_batch the one-warp K7).
mkdir -p gpurun_out
SKIP='not full_length_replay and not every_step_matches and not team_and_one_warp'
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 2400 compute-sanitizer --tool $tool $extra --error-exitcode 9 python -m pytest tests -m gpu -q -k "$SKIP" -p no:cacheprovider 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_sanitizer_$tool.txt
  echo "== $tool: exit ${PIPESTATUS[0]}"; tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
