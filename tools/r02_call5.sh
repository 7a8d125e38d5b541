#!/bin/sh
# Round-2 GPU call 5 (2 GPUs): gated tests, library NCCL paths, bench at N=1 and N=2.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_c5_pytest.log 2>&1
tail -15 gpurun_out/r02_c5_pytest.log
python tools/multi_gpu_check.py --group 2 > gpurun_out/r02_c5_group.log 2>&1
tail -3 gpurun_out/r02_c5_group.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r02_c5_sharded.log 2>&1
tail -4 gpurun_out/r02_c5_sharded.log
python bench.py > gpurun_out/r02_c5_bench_n1.json 2> gpurun_out/r02_c5_bench_n1.err
cut -c1-300 gpurun_out/r02_c5_bench_n1.json; tail -2 gpurun_out/r02_c5_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02_c5_bench_n2.json 2> gpurun_out/r02_c5_bench_n2.err
cut -c1-300 gpurun_out/r02_c5_bench_n2.json; tail -3 gpurun_out/r02_c5_bench_n2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c5_bench_ref.json 2> gpurun_out/r02_c5_bench_ref.err
cut -c1-400 gpurun_out/r02_c5_bench_ref.json
