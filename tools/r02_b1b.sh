#!/bin/bash
python -m pytest tests -m gpu -x -q -k "bit_identical or batch_invariance or full_size_healthy or slice_sampler or fused_groupnorm" 2>&1 | tail -3
for b in 1 2 4 8 16 64; do
for v in auto 0; do
MUDIFF_FUSED_GN=$v python bench.py --batch $b --steps 10 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=$b FUSED_GN=$v', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms')"
done; done
