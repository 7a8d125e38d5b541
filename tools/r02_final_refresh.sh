#!/bin/sh
# Refresh of the measurement half of tools/final_measure.sh with the last build (tests and sanitizers were run separately).
set -x
P=gpurun_out/r02_final
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1
SWEEP=final,kind python tools/r02_sweep.py > ${P}_sweep.log 2> ${P}_sweep.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
    --clock-control none -k regex:k_wf_trace -c 10 --csv --log-file ${P}_trace_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/traffic_from_ncu.py ${P}_trace_metrics.csv gpurun_out/traffic.json profiles/traffic.json
python bench.py > ${P}_bench.json 2> ${P}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in trace shade_a shade_b shade_c; do
  ncu --set full --clock-control none -k regex:k_wf_$k -s 1 -c 1 -f -o ${P}_prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py ${P}_prof_$k.ncu-rep > ${P}_ncu_k_wf_$k.txt 2>&1
  rm -f ${P}_prof_$k.ncu-rep
done
python tools/configs_bench.py > ${P}_configs.log 2> ${P}_configs.err
cut -c1-200 ${P}_bench.json
