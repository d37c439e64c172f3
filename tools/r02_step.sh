#!/bin/bash
# one GPU iteration of the round-2 kernel work: targeted tests, then A/B bench lines with per-kernel breakdowns
out=gpurun_out/${1:-r02b}
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "${2:-fused_groupnorm or discriminator or conv_tc}" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
for v in ${3:-0 2}; do
  MUDIFF_FUSED_GN=$v timeout 600 python bench.py --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --breakdown $out/breakdown_gn$v.txt > $out/bench_gn$v.json 2> $out/bench_gn$v.err; echo "bench gn=$v rc=$?"
  python -c "import json;d=json.load(open('$out/bench_gn$v.json'));print('GN=$v', round(d['value'],1),'slices/s', round(d['ms_per_step'],1),'ms  conv_tc', round(d['roofline']['kernel_ms_per_step'],1),'ms frac', round(d['roofline']['frac'],3))"
  head -8 $out/breakdown_gn$v.txt
done
