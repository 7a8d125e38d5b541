#!/bin/sh
# Round-2 GPU call 10: A/B of the software-pipelined node fetch (trace.pipe) on C4, bit-exactness checked per variant.
set -x
mkdir -p gpurun_out
SWEEP=pipe python tools/r02_sweep.py > gpurun_out/r02_c10_sweep_pipe.log 2> gpurun_out/r02_c10_sweep_pipe.err
cat gpurun_out/r02_c10_sweep_pipe.log; tail -5 gpurun_out/r02_c10_sweep_pipe.err
