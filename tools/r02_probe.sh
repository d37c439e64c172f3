#!/bin/bash
out=gpurun_out/${1:-r02p}
mkdir -p $out
Q="--no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e"
for b in 2 4 8 16 32; do
  timeout 300 python bench.py --batch $b --steps 5 $Q --no-roofline > $out/b$b.json 2> $out/b$b.err
  python -c "import json;d=json.load(open('$out/b$b.json'));print('B=$b', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms/step')"
done
for n in 64 128 256; do
  MUDIFF_FUSED_STATS_MIN_N=$n timeout 300 python bench.py $Q --breakdown $out/bd_fs$n.txt > $out/fs$n.json 2> $out/fs$n.err
  python -c "import json;d=json.load(open('$out/fs$n.json'));print('FUSED_STATS_MIN_N=$n', round(d['value'],1),'slices/s', round(d['ms_per_step'],1),'ms  conv_tc', round(d['roofline']['kernel_ms_per_step'],1))"
  head -5 $out/bd_fs$n.txt | tail -4
done
