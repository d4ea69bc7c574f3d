#!/bin/bash
# compute-sanitizer over the reduced parity subset (the reference builds its tests with UBSan,
# snappy.nimble:21,26-28).  memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the hash
# tables, the CRC tables, the v6 record rings); synccheck: invalid barrier / warp-sync use.  Summaries -> gpurun_out/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_$tool.log python scripts/sanitize_subset.py > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.out
  tail -3 gpurun_out/sanitize_$tool.log; tail -2 gpurun_out/sanitize_$tool.out
done
timeout 2400 $CS --tool racecheck --racecheck-report analysis --print-limit 20 --log-file gpurun_out/sanitize_racecheck.log python scripts/sanitize_subset.py small > gpurun_out/sanitize_racecheck.out 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.out
tail -3 gpurun_out/sanitize_racecheck.log; tail -2 gpurun_out/sanitize_racecheck.out
