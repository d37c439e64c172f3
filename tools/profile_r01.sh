#!/bin/bash
# ncu evidence for profiles/: launch list, DRAM bytes of every conv_tc launch of one step, --set full of the hot kernels.
# Run only after `python bench.py` exits 0 without ncu.  Numbers printed under ncu are not bench values.
# gpurun copies back at most 64 MiB: reports are exported to CSV on the box and dropped if large.
set -x
O=gpurun_out/prof
mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
$B > $O/plain.json 2> $O/plain.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3100 --csv --log-file $O/launches.csv $B > $O/ncu1.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -c 420 --csv --log-file $O/conv_tc_dram.csv $B > $O/ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"conv_tc_kernel|gn_apply_kernel|gn_stats_kernel|fir4_quad|conv_stem_gn|stem_moments" -c 30 -o $O/full_a -f $B > $O/ncu3.log 2>&1
ncu -i $O/full_a.ncu-rep --page raw --csv > $O/full_a_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel|conv_head_kernel" -c 2 -o $O/full_b -f $B > $O/ncu4.log 2>&1
ncu -i $O/full_b.ncu-rep --page raw --csv > $O/full_b_raw.csv 2>/dev/null
for f in $O/full_a.ncu-rep $O/full_b.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 20000000 ]; then rm -f $f; fi; done
gzip -f $O/launches.csv
du -sh $O; ls -la $O
