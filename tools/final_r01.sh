#!/bin/bash
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final/pytest.txt
timeout 600 python bench.py --breakdown gpurun_out/final/breakdown.txt > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final/bench_ref.json 2>/dev/null; echo "ref rc=$?"
cut -c1-200 gpurun_out/final/bench.json; cut -c1-200 gpurun_out/final/bench_ref.json
ncu --set full --clock-control none -k regex:"attn_tc_kernel" -c 1 -o gpurun_out/final/attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline > gpurun_out/final/ncu_attn.log 2>&1
ncu -i gpurun_out/final/attn.ncu-rep --page raw --csv > gpurun_out/final/attn_raw.csv 2>/dev/null
rm -f gpurun_out/final/attn.ncu-rep
python __graft_entry__.py smoke 2>&1 | tail -2
