#!/bin/sh
# Round-2 2-GPU evidence with the final build: library NCCL paths vs one GPU, and the N=2 bench line.
set -x
P=gpurun_out/r02_final_n2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > ${P}_render_sharded.log 2>&1
tail -3 ${P}_render_sharded.log
python tools/multi_gpu_check.py --group 2 > ${P}_group_render.log 2>&1
tail -3 ${P}_group_render.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > ${P}_bench.json 2> ${P}_bench.err
cut -c1-400 ${P}_bench.json; tail -2 ${P}_bench.err
