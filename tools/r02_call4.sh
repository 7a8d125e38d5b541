#!/bin/sh
# Round-2 GPU call 4: gated GPU tests on the bounded-MIS build, then the option sweep.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c4_pytest.log 2>&1
tail -15 gpurun_out/r02_c4_pytest.log
python tools/r02_sweep.py > gpurun_out/r02_c4_sweep.log 2> gpurun_out/r02_c4_sweep.err
cat gpurun_out/r02_c4_sweep.log
tail -5 gpurun_out/r02_c4_sweep.err
python bench.py > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err
cat gpurun_out/r02_c4_bench.json | cut -c1-600
