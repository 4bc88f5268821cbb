#!/usr/bin/env python
"""Decode kernels alone: stages two lz4-compressed C2 shards (31.25 M rows, 9375 value streams of 80 KB per shard over
the three columns the query reads) once with the eight-lanes-per-stream kernel and once with the lane-per-stream one.
Run it under `ncu --metrics gpu__time_duration.sum -k regex:cg_lz4_lane_kernel|cg_decompress_kernel --csv` and feed the
CSV to this script with --summarise to get decoded GB/s per launch."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        except ValueError:
            continue
        out.setdefault(r[ki].split("(")[0], []).append(v)
    decoded = 9375 * 80000
    for k, v in out.items():
        avg = sum(v) / len(v)
        print(f"{k}: {len(v)} launches, {avg:.1f} us average = {decoded / avg / 1e3:.1f} GB/s decoded (750 MB per launch)")
        print("   per launch (us):", " ".join(f"{x:.0f}" for x in v))
    sys.exit(0)

import bench  # noqa: E402
from citus_b200 import columnar as cg  # noqa: E402

cg.init(0)
cg.set_writer_compression("lz4")
rels = {s: cg.Relation.generate(list(bench.C2_COLUMNS), 31_250_000, seed=bench.SEED, first_row=s * 31_250_000,
                                stripe_row_limit=bench.STRIPE_ROWS, chunk_row_limit=bench.CHUNK_ROWS, nthreads=32) for s in range(2)}
cg.set_writer_compression("none")
# launch order (ten launches each): groups, lanes at <= 32 warps/SM (one stream per warp here), then 6 (4 streams per
# warp), 3 (8 per warp), 1 (32 per warp)
# groups; lanes in step (32 streams per warp); lanes on their own, one stream per warp
for mode, warps in ((0, 0), (1, 0), (1, 64)):
    cg.set_option("lz4_lanes", mode)
    cg.set_option("lz4_lane_warps", warps)
    for s in rels:
        sh = cg.Shard(rels[s], [0, 1, 2])
        sh.free()
print("done")
