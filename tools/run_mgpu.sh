#!/bin/bash
# usage: tools/run_mgpu.sh NGPU [extra args for mgpu_parity.py]   (run on the GPU box under gpurun --gpus N)
set -u
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/mgpu_${N}_smi.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/mgpu_parity.py "$@" > gpurun_out/mgpu_${N}.log 2>&1
echo "rc=$?" >> gpurun_out/mgpu_${N}.log
tail -30 gpurun_out/mgpu_${N}.log
