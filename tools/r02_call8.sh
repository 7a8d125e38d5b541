#!/bin/sh
# Round-2 GPU call 8: full gated suite (textures, worker, device update_frame) + keyframed-path variants on tr15.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c8_pytest.log 2>&1
tail -15 gpurun_out/r02_c8_pytest.log
python tools/c5_bench.py > gpurun_out/r02_c8_c5.log 2> gpurun_out/r02_c8_c5.err
cut -c1-330 gpurun_out/r02_c8_c5.log; tail -3 gpurun_out/r02_c8_c5.err
