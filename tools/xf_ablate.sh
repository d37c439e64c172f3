#!/bin/bash
for m in 0 1 2; do
  MUDIFF_FUSED_GN=1 MUDIFF_XF_DBG=$m timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --breakdown gpurun_out/bd_xf$m.txt > gpurun_out/b_xf$m.json 2>/dev/null
  echo "XF_DBG=$m"; head -3 gpurun_out/bd_xf$m.txt; grep -E "K= 2880|N=  64 K=  576|N= 128 K= 3456" gpurun_out/bd_xf$m.txt
done
