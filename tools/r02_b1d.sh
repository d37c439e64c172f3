#!/bin/bash
# Batch-1 timing ablations of the conv kernels inside the whole loop (results are wrong with these flags; timing only):
# how much of a batch-1 conv launch is operand traffic, how much epilogue, how much fixed cost?
Q="--steps 20 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline"
for b in 1 4; do
for f in 0 64 128 192; do
MUDIFF_CONV_DBG_FLAGS=$f python bench.py --batch $b $Q 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=$b flags=$f', round(d['ms_per_step'],2),'ms')"
done; done
