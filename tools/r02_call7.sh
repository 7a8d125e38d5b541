#!/bin/sh
# Round-2 GPU call 7: where the time goes in the keyframed (tr15) path; worker + PNG tests on the device.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_worker_and_png.py tests/test_abi_and_host.py -m gpu -x -q > gpurun_out/r02_c7_pytest.log 2>&1
tail -5 gpurun_out/r02_c7_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c7_launches_tr15.csv python tools/c5_bench.py --quick > gpurun_out/r02_c7_c5.log 2>&1
tail -2 gpurun_out/r02_c7_c5.log
