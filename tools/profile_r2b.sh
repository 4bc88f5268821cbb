#!/bin/bash
# Round-2 profiling pass B: the rebuilt scatter / routing kernels and the bulk-copy realign kernel (A/B against the LDG kernel).
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02b_launches_c4.csv \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > gpurun_out/r02b_launches_c4.json 2> gpurun_out/r02b_launches_c4.err
for tma in 1 0; do
  CG_REALIGN_TMA=$tma $NCU --metrics gpu__time_duration.sum -k regex:cg_realign -c 200 --csv --log-file gpurun_out/r02b_realign_tma$tma.csv \
      python bench.py --steps 1 --warmup 1 --no-extra --no-cpu --e2e-steps 5 > /dev/null 2> gpurun_out/r02b_realign_tma$tma.err
done
$NCU --set full --import-source on -k regex:cg_scatter_staged -s 2 -c 1 -o gpurun_out/r02b_scatter \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02b_full_scatter.err
$NCU --set full --import-source on -k regex:cg_route_hist -s 2 -c 1 -o gpurun_out/r02b_route_hist \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02b_full_hist.err
$NCU --set full --import-source on -k regex:cg_realign_tma -s 20 -c 1 -o gpurun_out/r02b_realign_tma \
    python bench.py --steps 1 --warmup 1 --no-extra --no-cpu --e2e-steps 5 > /dev/null 2> gpurun_out/r02b_full_realign.err
$NCU --set full --import-source on -k regex:cg_jit_scan -s 40 -c 1 -o gpurun_out/r02b_jit_nullable \
    python bench.py --workload c2null --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02b_full_null.err
ls -la gpurun_out | grep r02b
