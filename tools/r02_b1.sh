#!/bin/bash
Q="--batch ${1:-1} --steps 20 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline"
run() { env "$@" python bench.py $Q 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('$*', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms', d['launches_per_step'],'launches')"; }
run A=0
run MUDIFF_GN_SINGLE_PASS=1
run MUDIFF_FUSED_GN=2
run MUDIFF_FUSED_GN=2 MUDIFF_GN_SINGLE_PASS=1
run MUDIFF_FUSED_GN=1
run MUDIFF_FUSED_STATS_MIN_N=64
run MUDIFF_FUSED_STATS_MIN_N=64 MUDIFF_FUSED_GN=2
