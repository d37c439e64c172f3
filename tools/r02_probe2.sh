#!/bin/bash
out=gpurun_out/${1:-r02q}
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "conv_tc or groupnorm or gn_" > $out/pytest.log 2>&1; tail -3 $out/pytest.log
Q="--no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e"
for n in 100000 64 128; do
  MUDIFF_FUSED_STATS_MIN_N=$n timeout 300 python bench.py $Q --breakdown $out/bd_fs$n.txt > $out/fs$n.json 2> $out/fs$n.err
  python -c "import json;d=json.load(open('$out/fs$n.json'));print('FUSED_STATS_MIN_N=$n', round(d['value'],1),'slices/s', round(d['ms_per_step'],1),'ms  conv_tc', round(d['roofline']['kernel_ms_per_step'],1))"
  head -6 $out/bd_fs$n.txt | tail -5
done
