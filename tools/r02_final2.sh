#!/bin/bash
# Final round-2 evidence after the streaming sampler / launch-shape changes: GPU test-suite, smoke, default bench line with
# breakdown, then the ncu launch lists (B = 64 and B = 1) of the same command.  Logs -> gpurun_out/final2
out=gpurun_out/final2; mkdir -p $out
bash tools/r02_check.sh final2
export MUDIFF_WAIT_CYCLES=0
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-volume --no-reference-gpu --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches.csv $B > $out/ncu1.log 2>&1; echo "ncu b64 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches_b1.csv $B --batch 1 > $out/ncu2.log 2>&1; echo "ncu b1 rc=$?"
gzip -f $out/launches.csv $out/launches_b1.csv
ls -la $out
