#!/usr/bin/env bash
# host-buffer pipeline (e2e) sweep for the Sins headline: utterance chunk schedules x compute streams
set -u
mkdir -p gpurun_out
for cfg in "4,8,12,6,2:1" "4,8,12,6,2:2" "8,12,8,4:1" "8,12,8,4:2" "6,10,10,6:1" "10,12,8,2:1" "16,12,4:1" "12,12,8:1" "12,14,6:2" "16,16:1" "3,6,9,8,4,2:2"; do
  ch=${cfg%%:*}; st=${cfg##*:}
  out=$(timeout 100 python bench.py --e2e-only --steps 10 --warmup 3 --e2e-chunks $ch --e2e-streams $st 2>/dev/null | tail -1)
  echo "$cfg $out" | tee -a gpurun_out/e2e_sweep.txt
done
