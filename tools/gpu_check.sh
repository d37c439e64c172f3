#!/bin/bash
# standard GPU check: parity tests, then the default bench with a per-kernel breakdown
tag=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_$tag.txt > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
head -22 gpurun_out/breakdown_$tag.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print({k:d[k] for k in ('value','ms_per_step','clocks')}, d['e2e']['value'], d['roofline']['achieved'])"
