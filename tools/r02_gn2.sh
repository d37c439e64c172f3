#!/bin/bash
# Statistics kernel with pipelined 64 KB chunks (cpb chunks per block) vs the old 256 KB / 32-chunk geometry, B = 64 and B = 1,
# plus the streaming end-to-end line.  tools/r02_gn2.sh <outdir>
out=gpurun_out/${1:-gn2}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k 'groupnorm or gn or streaming or invarian' > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
Q="--no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs"
i=0
for cfg in "A=0" "MUDIFF_GN_CHUNK_KB=256" "MUDIFF_GN_CPB=8" "MUDIFF_GN_CHUNK_KB=128"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py $Q --breakdown $out/breakdown_$i.txt > $out/bench_$i.json 2> $out/bench_$i.err; echo "[$cfg] rc=$?"
  python -c "import json;d=json.load(open('$out/bench_$i.json'));print('  ', round(d['value'],1),'slices/s', round(d['ms_per_step'],1),'ms  e2e', round(d['e2e']['value'],1), ' conv_tc', round(d['roofline']['kernel_ms_per_step'],1),'ms frac', round(d['roofline']['frac'],3), d['clocks'])"
  head -9 $out/breakdown_$i.txt | tail -8
done
Q1="--batch 1 --steps 20 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline"
for cfg in "A=0" "MUDIFF_GN_CHUNK_KB=256" "MUDIFF_GN_CHUNK_KB=32" "MUDIFF_GN_CHUNK_KB=128"; do
  env $cfg python bench.py $Q1 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=1 [$cfg]', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms', d['launches_per_step'],'launches')"
done
