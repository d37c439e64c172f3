#!/bin/bash
for v in 0 1 0 1; do
  MUDIFF_BCAP12=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_bcap$v.txt 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('BCAP12=$v', round(d['value'], 1), 'slices/s', round(d['roofline']['kernel_ms_per_step'], 1), 'ms conv_tc')"
done
for k in "K= 2880" "K= 2304" "K= 1728" "N=  64 K= 1152"; do grep -h "N=  64.*$k\|$k" gpurun_out/bd_bcap0.txt | head -1; grep -h "N=  64.*$k\|$k" gpurun_out/bd_bcap1.txt | head -1; done
