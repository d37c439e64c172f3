#!/bin/bash
# BASELINE configs[4]-style sanity sweep: other resolutions / widths through the same code path (short runs)
for cfg in "--size 128 --batch 64" "--size 512 --batch 8" "--size 256 --batch 16 --nf 128" "--size 256 --batch 1"; do
  echo "=== $cfg"
  timeout 300 python bench.py $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e 2> gpurun_out/sweep.err | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step', 'launches_per_step')}, d['config']['global_batch'], d['config']['size'])
    except Exception as e:
        print('PARSE', l[:200])
"
  tail -2 gpurun_out/sweep.err | cut -c1-300
done
