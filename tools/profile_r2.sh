#!/bin/bash
# Round-2 profiling pass on one B200 (run through gpurun): launch lists with per-launch device time and
# full ncu captures of the dominant kernels.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the headline command (no extras: the legs have their own lists)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --e2e-steps 5 > gpurun_out/r02_launches_bench.json 2> gpurun_out/r02_launches_bench.err
# 2. launch lists of the legs
for leg in c2null c4 c5 c1; do
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_$leg.csv \
      python bench.py --workload $leg --leg-steps 2 --no-cpu > gpurun_out/r02_launches_$leg.json 2> gpurun_out/r02_launches_$leg.err
done
# 3. full captures: one launch each of the kernels the numbers rest on
$NCU --set full --import-source on -k regex:cg_scan_fast_kernel -s 40 -c 1 -o gpurun_out/r02_scan_fast \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu --no-e2e > /dev/null 2> gpurun_out/r02_full_fast.err
$NCU --set full --import-source on -k regex:cg_jit_scan -s 40 -c 1 -o gpurun_out/r02_jit_nullable \
    python bench.py --workload c2null --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02_full_null.err
$NCU --set full --import-source on -k regex:cg_scatter_staged -s 2 -c 1 -o gpurun_out/r02_scatter \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02_full_scatter.err
$NCU --set full --import-source on -k regex:cg_route_hist -s 2 -c 1 -o gpurun_out/r02_route_hist \
    python bench.py --workload c4 --leg-steps 2 --no-cpu > /dev/null 2> gpurun_out/r02_full_hist.err
ls -la gpurun_out | tail -20
