#!/bin/bash
# round 2, final pass on one GPU: the whole GPU test suite, the default bench line, launch lists of the headline and of the
# TPC-H leg (paired 32-bit cells in the Q1 kernel), one full capture of the Q1 kernel
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log
timeout 400 python bench.py > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.log; tail -2 gpurun_out/r2m_bench_1gpu.log
NCU="ncu --clock-control none"
timeout 200 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2m_launches.csv \
    python bench.py --no-extra --no-e2e --no-cpu --steps 2 --warmup 1 > /dev/null 2> gpurun_out/r2m_launches.err
timeout 200 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r2m_launches_c5.csv \
    python bench.py --workload c5 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r2m_launches_c5.err
timeout 200 $NCU --set full --import-source on -k regex:cg_jit -s 240 -c 1 -o gpurun_out/r2m_jit_q1 \
    python bench.py --workload c5 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r2m_full_q1.err
ls -la gpurun_out | grep r2m
