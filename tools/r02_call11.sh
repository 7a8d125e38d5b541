#!/bin/sh
# Round-2 GPU call 11: gated suite on the new default trace variant (trace.pipe 36), bench, ncu per-instruction table.
set -x
P=gpurun_out/r02_c11
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1
tail -5 ${P}_pytest.log
python bench.py --no-cpu-baseline > ${P}_bench.json 2> ${P}_bench.err
cut -c1-600 ${P}_bench.json
ncu --set full --import-source on --clock-control none -k regex:k_wf_trace -s 1 -c 1 -f -o ${P}_prof_trace python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py ${P}_prof_trace.ncu-rep > ${P}_ncu_k_wf_trace.txt 2>&1
ncu -i ${P}_prof_trace.ncu-rep --page source --csv > ${P}_source_k_wf_trace.csv 2> /dev/null
rm -f ${P}_prof_trace.ncu-rep
