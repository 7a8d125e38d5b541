#!/bin/sh
# Round-2 GPU call 6: gated tests (device update_frame, per-path transform table), C5 timing, bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c6_pytest.log 2>&1
tail -15 gpurun_out/r02_c6_pytest.log
python tools/c5_bench.py > gpurun_out/r02_c6_c5.log 2> gpurun_out/r02_c6_c5.err
cat gpurun_out/r02_c6_c5.log; tail -3 gpurun_out/r02_c6_c5.err
python bench.py > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err
cut -c1-300 gpurun_out/r02_c6_bench.json
