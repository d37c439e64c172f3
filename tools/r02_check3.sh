#!/bin/bash
# GPU test-suite, smoke, default bench line with breakdown.  Logs -> gpurun_out/$1
out=gpurun_out/${1:-final3}; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest.log; tail -3 $out/pytest.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log; tail -3 $out/smoke.log
timeout 900 python bench.py --breakdown $out/breakdown.txt > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$out/bench.json'))
print(round(d['value'],1),'slices/s e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'volume',round(d['volume']['value'],1),d['clocks'])
for c in d['other_configs']: print('  ',c['config'],round(c['slices_per_s'],1),round(c['ms_per_step'],2))"
head -8 $out/breakdown.txt
