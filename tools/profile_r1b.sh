#!/bin/bash
# round-1 record: tools/tpch_bench.py named below was that round's driver; `bench.py --workload c5` replaced it
# round-1 (second session) evidence run on one GPU: TPC-H SF100 through the plan-specialised kernels,
# ncu launch lists, and full captures of the NVRTC kernels (Q1, Q6) and of the decompression kernel.
set -x
mkdir -p gpurun_out
python tools/tpch_bench.py --reps 3 2> gpurun_out/tpch_jit.err > gpurun_out/tpch_jit.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_tpch_jit.csv \
    python tools/tpch_bench.py --reps 1 > /dev/null 2> gpurun_out/ncu_tpch.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_lz4.csv \
    python bench.py --compression lz4 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_lz4.err
# Q6 launches come first (1 parity + 2 x 32), then Q1
ncu --set full --clock-control none --import-source on -k regex:cg_jit_scan -s 80 -c 1 -f -o gpurun_out/prof_jit_q1 \
    python tools/tpch_bench.py --reps 1 > /dev/null 2> gpurun_out/ncu_q1.err
ncu --set full --clock-control none --import-source on -k regex:cg_jit_scan -s 5 -c 1 -f -o gpurun_out/prof_jit_q6 \
    python tools/tpch_bench.py --reps 1 > /dev/null 2> gpurun_out/ncu_q6.err
ncu --set full --clock-control none --import-source on -k regex:cg_decompress -s 40 -c 1 -f -o gpurun_out/prof_decompress \
    python bench.py --compression lz4 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_dec.err
ls -la gpurun_out/*.ncu-rep
