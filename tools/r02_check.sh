#!/bin/bash
# Round-2 GPU check: GPU test-suite, smoke, default bench line, reference GPU baseline table.  Logs -> gpurun_out/$1
out=gpurun_out/${1:-r02}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt
nproc >> $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log
tail -3 $out/smoke.log
timeout 900 python bench.py --breakdown $out/breakdown.txt > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json
timeout 900 python tools/ref_gpu_baseline.py --out $out/reference_gpu.json > $out/reference_gpu.log 2>&1; echo "refgpu rc=$?"
grep -v "^\[" $out/reference_gpu.log | tail -8
