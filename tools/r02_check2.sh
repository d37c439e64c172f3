#!/bin/bash
# Round-2 GPU check without the reference-GPU table: GPU test-suite, smoke, default bench line + breakdown.  Logs -> gpurun_out/$1
out=gpurun_out/${1:-r02c}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt
nproc >> $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log
tail -3 $out/smoke.log
timeout 900 python bench.py --breakdown $out/breakdown.txt > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json
