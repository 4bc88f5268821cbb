#!/bin/bash
# Prints the GPU box's host/GPU facts that DESIGN.md and bench.py sizing rely on.
mkdir -p gpurun_out
{
echo "== nproc"; nproc
echo "== mem"; free -g
echo "== cpu"; lscpu | head -25
echo "== nvidia-smi"; nvidia-smi
echo "== topo"; nvidia-smi topo -m
echo "== nccl"; ls /usr/include/nccl.h /usr/lib/x86_64-linux-gnu/libnccl* 2>&1
echo "== lz4/zstd"; ls /usr/lib/x86_64-linux-gnu/liblz4* /usr/lib/x86_64-linux-gnu/libzstd* /usr/include/lz4.h /usr/include/zstd.h 2>&1
echo "== df"; df -h /tmp /dev/shm . 2>&1
echo "== ulimit"; ulimit -a
} > gpurun_out/probe_box.txt 2>&1
cat gpurun_out/probe_box.txt | head -80
