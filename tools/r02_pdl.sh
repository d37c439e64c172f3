#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 1 4 64; do
for v in 1 0; do
MUDIFF_PDL=$v python bench.py --batch $b --steps 10 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=$b PDL=$v', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms')"
done; done
