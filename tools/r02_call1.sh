#!/bin/sh
# Round-2 GPU call 1: gated GPU tests (incl. the new full-size one-call parity file), smoke, bench line, fresh ncu captures
# of the three wavefront kernels with L1/L2 metrics, and compute-sanitizer over both film kernels.
#   gpurun --timeout 2400 -- 'sh tools/r02_call1.sh'
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c1_pytest.log 2>&1
tail -5 gpurun_out/r02_c1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c1_smoke.log 2>&1
python bench.py > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in trace shade film_v2; do
  ncu --set full --import-source on --clock-control none -k regex:k_wf_$k -s 1 -c 1 -f -o gpurun_out/r02_c1_prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r02_c1_prof_$k.ncu-rep > gpurun_out/r02_c1_ncu_k_wf_$k.txt 2>&1
done
for t in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $t python tools/sanitize.py > gpurun_out/r02_c1_sanitizer_$t.log 2>&1
  tail -4 gpurun_out/r02_c1_sanitizer_$t.log
done
ls -la gpurun_out | tail -20
