#!/bin/bash
# A/B at a given batch: tools/r02_ab_b.sh <outdir> <batch> "VAR=val ..." ...
out=gpurun_out/${1:-abb}; shift
bs=$1; shift
mkdir -p $out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --batch $bs --steps 10 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --breakdown $out/breakdown_b${bs}_$i.txt > $out/bench_b${bs}_$i.json 2> $out/bench_b${bs}_$i.err; echo "[B=$bs $cfg] rc=$?"
  python -c "import json;d=json.load(open('$out/bench_b${bs}_$i.json'));print('  ', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms launches', d['launches_per_step'])"
  head -6 $out/breakdown_b${bs}_$i.txt | tail -5
done
