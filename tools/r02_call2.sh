#!/bin/sh
# Round-2 GPU call 2: gated GPU tests on the sort + split-shade build, then the option sweep.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c2_pytest.log 2>&1
tail -15 gpurun_out/r02_c2_pytest.log
python tools/r02_sweep.py > gpurun_out/r02_c2_sweep.log 2> gpurun_out/r02_c2_sweep.err
cat gpurun_out/r02_c2_sweep.log
tail -5 gpurun_out/r02_c2_sweep.err
