#!/bin/bash
# Batch-1 check of the launch-shape changes (stem rows per block, head strip height) + an ncu --set full capture of a few
# batch-1 conv_tc launches with source-level stall sampling.  tools/r02_b1prof.sh <outdir>
out=gpurun_out/${1:-b1prof}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
Q1="--batch 1 --steps 20 --no-cpu-baseline --no-volume --no-reference-gpu --no-other-configs --no-e2e --no-roofline"
python bench.py $Q1 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=1', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms', d['launches_per_step'],'launches')"
python bench.py --batch 4 --steps 10 ${Q1#--batch 1 --steps 20} 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().splitlines()[-1]);print('B=4', round(d['value'],1),'slices/s', round(d['ms_per_step'],2),'ms')"
export MUDIFF_WAIT_CYCLES=0
B1="python bench.py --batch 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-volume --no-reference-gpu --no-other-configs"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip 700 --launch-count 14 -o $out/conv_b1 -f $B1 > $out/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $out
