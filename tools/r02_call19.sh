#!/bin/sh
# Round-2 GPU call 19: tr15.json — keyframed trace variants at 7 / 8 / 9 CTAs per SM; ncu per-instruction profile of k_wf_shade_b.
set -x
P=gpurun_out/r02_c19
python tools/c5_bench.py --quick > ${P}_c5.log 2> ${P}_c5.err
cut -c1-330 ${P}_c5.log
ncu --set full --import-source on --clock-control none -k regex:k_wf_shade_b -s 12 -c 1 -f -o ${P}_prof_shade_b python tools/c5_bench.py --quick > /dev/null 2>&1
python tools/ncu_summary.py ${P}_prof_shade_b.ncu-rep > ${P}_ncu_k_wf_shade_b.txt 2>&1
ncu -i ${P}_prof_shade_b.ncu-rep --page source --csv > ${P}_source_k_wf_shade_b.csv 2> /dev/null
rm -f ${P}_prof_shade_b.ncu-rep
head -12 ${P}_ncu_k_wf_shade_b.txt
