#!/usr/bin/env python
"""Decode-path probe: end-to-end scan of compressed C2 shards from pinned pages (DMA path), a few full-size
shards, per codec.  CG_TRACE=1 adds the per-shard kernel times (realign, decode + scan)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from citus_b200 import columnar as cg  # noqa: E402

codec = sys.argv[1] if len(sys.argv) > 1 else "lz4"
nshards = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rows = 31_250_000
cg.init(0)
cg.numa_bind()
torch.cuda.set_stream(torch.cuda.Stream())
cg.use_torch_stream()
cg.set_writer_compression(codec)
rels = bench.generate_shards(cg, list(range(nshards)), rows, 32)
cg.set_writer_compression("none")
aggs = [cg.sum_(2), cg.count_star()]
desc = cg.make_desc(bench.QUALS, bench.GROUP, aggs)
aggs[0].term_abs_bound = max(cg.relation_bounds(r, desc)[2][0] for r in rels.values())
desc = cg.make_desc(bench.QUALS, bench.GROUP, aggs)
partial = cg.GpuColumnarAgg(desc, rels[0].column_descs(), 0, bench.NKEYS - 1, rows * nshards)
for r in rels.values():
    r.register()
for rep in range(3):
    partial.reset()
    torch.cuda.synchronize()
    t0 = time.time()
    h2d = 0
    for s in rels:
        st = partial.scan_relation(rels[s], want_stats=(rep == 0))
        h2d += st.h2d_bytes
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"{codec} rep {rep}: {nshards} shards in {dt * 1e3:.1f} ms = {nshards * rows / dt / 1e9:.2f} G rows/s, h2d {h2d / 1e9:.2f} GB", flush=True)
