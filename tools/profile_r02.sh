#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of the default bench command, DRAM bytes of every conv_tc launch of one
# step, --set full of the hot kernels.  Run only after `python bench.py` exits 0 without ncu; numbers printed under ncu are
# not bench values.  MUDIFF_WAIT_CYCLES=0: ncu replays a kernel ~40 times, the kernels' mbarrier-wait bound must not trip.
set -x
O=gpurun_out/prof2
mkdir -p $O
export MUDIFF_WAIT_CYCLES=0
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-volume --no-reference-gpu --no-other-configs"
$B > $O/plain.json 2> $O/plain.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv $B > $O/ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc -c 420 --csv --log-file $O/conv_tc_dram.csv $B > $O/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_tc|gn_apply_kernel|gn_stats_kernel|fir4_quad|conv_stem_gn|conv_head|attn_tc" -c 40 -o $O/full_a -f $B > $O/ncu3.log 2>&1
ncu -i $O/full_a.ncu-rep --page raw --csv > $O/full_a_raw.csv 2>/dev/null
B1="python bench.py --batch 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-volume --no-reference-gpu --no-other-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_b1.csv $B1 > $O/ncu5.log 2>&1
for f in $O/full_a.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 30000000 ]; then rm -f $f; fi; done
gzip -f $O/launches.csv $O/launches_b1.csv $O/conv_tc_dram.csv
du -sh $O; ls -la $O
