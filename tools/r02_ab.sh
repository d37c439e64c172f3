#!/bin/bash
# A/B bench lines under different environment switches: tools/r02_ab.sh <outdir> "VAR=val VAR2=val" "VAR=val" ...
out=gpurun_out/${1:-ab}; shift
mkdir -p $out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --breakdown $out/breakdown_$i.txt > $out/bench_$i.json 2> $out/bench_$i.err; echo "[$cfg] rc=$?"
  python -c "import json;d=json.load(open('$out/bench_$i.json'));print('  ', round(d['value'],1),'slices/s', round(d['ms_per_step'],1),'ms  conv_tc', round(d['roofline']['kernel_ms_per_step'],1),'ms frac', round(d['roofline']['frac'],3), d['clocks'])"
  head -9 $out/breakdown_$i.txt | tail -8
done
