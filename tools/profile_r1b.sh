#!/bin/bash
# round-1 (second session) evidence run: parity tests, the lz4 variant of the bench, ncu launch lists and
# full captures of the plan-specialised (NVRTC) kernel and the decompression kernel.  One GPU.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --compression lz4 --no-cpu > gpurun_out/bench_lz4.json 2> gpurun_out/bench_lz4.err
tail -4 gpurun_out/bench_lz4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tpch_jit.csv \
    python tools/tpch_bench.py --rows 60000000 --reps 1 > /dev/null 2> gpurun_out/ncu_tpch.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_lz4.csv \
    python bench.py --rows 125000000 --compression lz4 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_lz4.err
ncu --set full --clock-control none --import-source on -k regex:cg_jit_scan -s 70 -c 1 -f -o gpurun_out/prof_jit_q1 \
    python tools/tpch_bench.py --rows 60000000 --reps 1 > /dev/null 2> gpurun_out/ncu_q1.err
ncu --set full --clock-control none --import-source on -k regex:cg_jit_scan -s 5 -c 1 -f -o gpurun_out/prof_jit_q6 \
    python tools/tpch_bench.py --rows 60000000 --reps 1 > /dev/null 2> gpurun_out/ncu_q6.err
ncu --set full --clock-control none --import-source on -k regex:cg_decompress -s 8 -c 1 -f -o gpurun_out/prof_decompress \
    python bench.py --rows 125000000 --compression lz4 --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > /dev/null 2> gpurun_out/ncu_dec.err
ls -la gpurun_out/*.ncu-rep
