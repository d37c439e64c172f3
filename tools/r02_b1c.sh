#!/bin/bash
out=gpurun_out/${1:-b1c}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
Q="--no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline"
for b in 1 2 4; do
python bench.py --batch $b --steps 20 $Q 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=$b', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms', d['launches_per_step'],'launches')"
done
