#!/bin/sh
# Round-2 GPU call 15: per-CUDA-source-line profile of the fused shade kernel (C4, round 1).
set -x
P=gpurun_out/r02_c15
ncu --set full --import-source on --clock-control none -k regex:k_wf_shade -s 1 -c 1 -f -o ${P}_prof_shade python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i ${P}_prof_shade.ncu-rep --page source --print-source cuda --csv > ${P}_cuda_k_wf_shade.csv 2> ${P}_cuda.err
head -c 600 ${P}_cuda_k_wf_shade.csv; tail -2 ${P}_cuda.err
python tools/ncu_summary.py ${P}_prof_shade.ncu-rep > ${P}_ncu_k_wf_shade.txt 2>&1
rm -f ${P}_prof_shade.ncu-rep
