#!/bin/sh
# End-of-round evidence in one GPU-box call (outputs under gpurun_out/, copied to profiles/ afterwards):
#   gpurun --timeout 1500 -- 'sh tools/final_measure.sh'
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_wf_trace -c 10 --csv --log-file gpurun_out/final_trace_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/traffic_from_ncu.py gpurun_out/final_trace_dram.csv gpurun_out/traffic.json profiles/traffic.json
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python tools/c5_bench.py > gpurun_out/final_c5.json 2> gpurun_out/final_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_wf_trace -s 1 -c 1 -f -o gpurun_out/prof_trace_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_wf_shade -s 1 -c 1 -f -o gpurun_out/prof_shade_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -14
