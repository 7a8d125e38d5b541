#!/bin/sh
# Round-2 GPU call 9 (after the container was re-created): gated suite, smoke, bench (both arms), launch list,
# ncu --set full of the trace and shade kernels with the per-instruction source table.
set -x
P=gpurun_out/r02_c9
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1
tail -4 ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1
tail -2 ${P}_smoke.log
python bench.py > ${P}_bench.json 2> ${P}_bench.err
cat ${P}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err
cat ${P}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in trace shade; do
  ncu --set full --import-source on --clock-control none -k regex:k_wf_$k -s 1 -c 1 -f -o ${P}_prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py ${P}_prof_$k.ncu-rep > ${P}_ncu_k_wf_$k.txt 2>&1
  ncu -i ${P}_prof_$k.ncu-rep --page source --csv > ${P}_source_k_wf_$k.csv 2> /dev/null
  rm -f ${P}_prof_$k.ncu-rep
done
ls -la gpurun_out | tail -15
