#!/bin/bash
# stem_moments_scope: GPU tests + batch-1 / batch-64 bench lines
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
Q="--no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-roofline"
python bench.py --batch 1 --steps 20 $Q --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=1', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms', d['launches_per_step'],'launches')"
python bench.py $Q 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=64', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],1), d['launches_per_step'],'launches')"
