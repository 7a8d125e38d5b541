#!/bin/sh
# Round-2 GPU call 3: gated GPU tests; per-kernel times and ncu of the split shade kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c3_pytest.log 2>&1
tail -15 gpurun_out/r02_c3_pytest.log
export TRB_SORT=0 TRB_SHADE_SPLIT=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_c3_launches_split.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for k in shade_a shade_b shade_c; do
  ncu --set full --import-source on --clock-control none -k regex:k_wf_$k -s 1 -c 1 -f -o gpurun_out/r02_c3_prof_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r02_c3_prof_$k.ncu-rep > gpurun_out/r02_c3_ncu_k_wf_$k.txt 2>&1
done
ls -la gpurun_out | tail -8
