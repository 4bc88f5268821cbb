#!/usr/bin/env python
"""Where the end-to-end (host pages -> result) time goes: host time of every cg_scan_relation
call, of the fetch, and the device time of the whole pass, for a few full-size C2 shards."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from citus_b200 import columnar as cg  # noqa: E402

nshards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows = 31_250_000
cg.init(0)
print("numa node", cg.numa_bind())
torch.cuda.set_stream(torch.cuda.Stream())
cg.use_torch_stream()
rels = {s: cg.Relation.generate(list(bench.C2_COLUMNS), rows, seed=bench.SEED, first_row=s * rows, stripe_row_limit=bench.STRIPE_ROWS,
                                chunk_row_limit=bench.CHUNK_ROWS, nthreads=32) for s in range(nshards)}
aggs = [cg.sum_(2), cg.count_star()]
desc = cg.make_desc(bench.C2_QUALS, bench.C2_GROUP, aggs)
aggs[0].term_abs_bound = max(cg.relation_bounds(r, desc)[2][0] for r in rels.values())
desc = cg.make_desc(bench.C2_QUALS, bench.C2_GROUP, aggs)
partial = cg.GpuColumnarAgg(desc, rels[0].column_descs(), 0, bench.NKEYS - 1, rows * nshards)
for r in rels.values():
    r.register()
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    partial.reset()
    calls = []
    for s in rels:
        t = time.time()
        partial.scan_relation(rels[s], want_stats=False)
        calls.append((time.time() - t) * 1e3)
    t_issue = time.time() - t0
    t = time.time()
    torch.cuda.synchronize()
    t_wait = time.time() - t
    t = time.time()
    res = partial.fetch(reuse=True)
    t_fetch = time.time() - t
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: calls ms {[round(c, 2) for c in calls]}  issue {t_issue * 1e3:.1f} ms, drain {t_wait * 1e3:.1f} ms, "
          f"fetch {t_fetch * 1e3:.1f} ms, device total {e0.elapsed_time(e1):.1f} ms, groups {res['n']}")
