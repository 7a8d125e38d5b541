#!/bin/sh
# End-of-round evidence in one GPU-box call (outputs under gpurun_out/, copied to profiles/ afterwards):
#   gpurun --timeout 2400 -- 'sh tools/final_measure.sh'
set -x
P=gpurun_out/r02_final
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1
tail -4 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1
SWEEP=final python tools/r02_sweep.py > ${P}_sweep.log 2> ${P}_sweep.err
# what ncu measures for the trace launches of ONE bench step: DRAM bytes, L2 sectors, L1 global-load sectors, L1TEX data pipe, issue slots
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum \
    --clock-control none -k regex:k_wf_trace -c 10 --csv --log-file ${P}_trace_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/traffic_from_ncu.py ${P}_trace_metrics.csv gpurun_out/traffic.json profiles/traffic.json
python bench.py > ${P}_bench.json 2> ${P}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in trace shade film_v2; do
  ncu --set full --import-source on --clock-control none -k regex:k_wf_$k -s 1 -c 1 -f -o ${P}_prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py ${P}_prof_$k.ncu-rep > ${P}_ncu_k_wf_$k.txt 2>&1
  if [ $k = trace ]; then ncu -i ${P}_prof_$k.ncu-rep --page source --csv > ${P}_source_k_wf_$k.csv 2> /dev/null; fi
  rm -f ${P}_prof_$k.ncu-rep   # gpurun merges at most 64 MiB back: keep the summaries and the per-instruction table, not the report
done
for t in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $t python tools/sanitize.py > ${P}_sanitizer_$t.log 2>&1
done
python tools/c5_bench.py > ${P}_c5.log 2> ${P}_c5.err
python tools/configs_bench.py > ${P}_configs.log 2> ${P}_configs.err
ls -la gpurun_out | tail -25
