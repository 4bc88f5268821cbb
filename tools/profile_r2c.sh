#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r02c_launches_c4.csv \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > gpurun_out/r02c_launches_c4.json 2> gpurun_out/r02c_launches_c4.err
$NCU --set full --import-source on -k regex:cg_scatter_staged -s 2 -c 1 -o gpurun_out/r02c_scatter \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02c_full_scatter.err
$NCU --set full --import-source on -k regex:cg_jit_scan -s 40 -c 1 -o gpurun_out/r02c_jit_nullable \
    python bench.py --workload c2null --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02c_full_null.err
ls -la gpurun_out | grep r02c
