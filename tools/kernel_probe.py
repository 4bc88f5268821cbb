#!/usr/bin/env python
"""Kernel-level probe: runs several query shapes over HBM-resident synthetic shards and
prints the fused kernel's device time and algorithmic GB/s for each (CUDA events inside
the library).  Used to find which part of the kernel bounds the rate:

    python tools/kernel_probe.py [--rows 62500000] [--shards 2] [--reps 5] [--shapes a,b,...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from citus_b200 import build  # noqa: E402

build.build()
from citus_b200 import columnar as cg  # noqa: E402

NKEYS = 1_000_000
COLUMNS = [(8, 0, 0, NKEYS, 0), (8, 0, 0, 100, 0), (8, 0, -10**9, 10**9, 0)] + [(8, 0, 0, 1 << 40, 0)] * 5


def shapes():
    s = {}
    s["scan_only_sum"] = dict(quals=[(1, "<", 50)], group=[], aggs=[cg.sum_(2), cg.count_star()])
    s["scan_nofilter_sum"] = dict(quals=[], group=[], aggs=[cg.sum_(2), cg.sum_(0), cg.sum_(1)])
    s["dense_count"] = dict(quals=[(1, "<", 50)], group=[0], aggs=[cg.count_star()])
    s["dense_sum_count"] = dict(quals=[(1, "<", 50)], group=[0], aggs=[cg.sum_(2), cg.count_star()])
    s["dense_sum_count_all"] = dict(quals=[], group=[0], aggs=[cg.sum_(2), cg.count_star()])
    s["dense_sum_count_sel10"] = dict(quals=[(1, "<", 10)], group=[0], aggs=[cg.sum_(2), cg.count_star()])
    s["hash_sum_count"] = dict(quals=[(1, "<", 50)], group=[0], aggs=[cg.sum_(2), cg.count_star()], hash=True)
    s["dense_2limb"] = dict(quals=[(1, "<", 50)], group=[0], aggs=[cg.sum_(2), cg.count_star()], nobound=True)
    return s


# synthetic lineitem (BASELINE config 5): decimals as scaled int64 cents, flags as 1-byte ints, dates int4
LINEITEM = [(8, 0, 100, 5100, 0),          # 0 l_quantity      1.00 .. 50.99
            (8, 0, 90000, 10500000, 0),    # 1 l_extendedprice
            (8, 0, 0, 11, 0),              # 2 l_discount      0.00 .. 0.10
            (8, 0, 0, 9, 0),               # 3 l_tax
            (1, 0, 65, 68, 0),             # 4 l_returnflag    3 values
            (1, 0, 70, 72, 0),             # 5 l_linestatus    2 values
            (4, 0, -2922, -365, 0),        # 6 l_shipdate      1992-01-01 .. 1998-12-31 (days since 2000-01-01)
            (4, 0, 1, 10001, 0)]           # 7 l_suppkey


def lineitem_shapes():
    s = {}
    s["tpch_q6"] = dict(quals=[(6, ">=", -2192), (6, "<", -1827), (2, ">=", 5), (2, "<=", 7), (0, "<", 2400)],
                        group=[], aggs=[cg.Agg(2, [(1, 0, 1), (2, 0, 1)])])
    s["tpch_q1"] = dict(quals=[(6, "<=", -486)], group=[4, 5],
                        aggs=[cg.sum_(0), cg.sum_(1), cg.Agg(2, [(1, 0, 1), (2, 100, -1)]),
                              cg.Agg(2, [(1, 0, 1), (2, 100, -1), (3, 100, 1)]), cg.sum_(2), cg.count_star()])
    return s


def run_shapes(all_shapes, rels, shards, a, want, out, expected_groups):
    for name, sh in all_shapes.items():
        if want and name not in want:
            continue
        d = cg.make_desc(sh["quals"], sh["group"], sh["aggs"], expected_groups=expected_groups)
        kmin, kmax, bounds, rows = cg.relation_bounds(rels[0], d)
        for r in rels[1:]:
            a2, b2, bd2, _ = cg.relation_bounds(r, d)
            kmin, kmax = min(kmin, a2), max(kmax, b2)
            bounds = [max(x, y) for x, y in zip(bounds, bd2)]
        for i, ag in enumerate(sh["aggs"]):
            ag.term_abs_bound = 0 if sh.get("nobound") else bounds[i]
        d = cg.make_desc(sh["quals"], sh["group"], sh["aggs"], expected_groups=expected_groups)
        if sh.get("hash"):
            kmin, kmax = 0, -1
        agg = cg.GpuColumnarAgg(d, rels[0].column_descs(), kmin, kmax, a.rows)
        best = None
        tot_bytes = 0
        for rep in range(a.reps):
            agg.reset()
            ms = 0.0
            tot_bytes = 0
            for shd in shards:
                st = agg.scan_shard(shd)
                ms += st.kernel_ms
                tot_bytes += st.bytes_scanned
            best = ms if best is None else min(best, ms)
        gbs = tot_bytes / best / 1e6
        out[name] = dict(ms=best, gbs=gbs, rows_per_s=a.rows / best * 1e3, groups=agg.ngroups())
        print(f"{name:24s} {best:8.3f} ms  {gbs:8.1f} GB/s  {a.rows / best / 1e6:8.2f} Grows/s  groups={out[name]['groups']}",
              flush=True)
        agg.free()


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rows", type=int, default=62_500_000)
    p.add_argument("--shards", type=int, default=2)
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--shapes", default="")
    p.add_argument("--null-ppm", type=int, default=0)
    a = p.parse_args()
    cg.init(0)
    cols = list(COLUMNS)
    if a.null_ppm:
        cols[2] = (8, 0, -10**9, 10**9, a.null_ppm)
    per = a.rows // a.shards
    rels = [cg.Relation.generate(cols, per, seed=20260922, first_row=s * per, nthreads=32) for s in range(a.shards)]
    shards = [cg.Shard(r, [0, 1, 2]) for r in rels]
    want = [x for x in a.shapes.split(",") if x]
    out = {}
    run_shapes(shapes(), rels, shards, a, want, out, NKEYS)
    if not want or any(w.startswith("tpch") for w in want):
        for sh in shards:
            sh.free()
        rels = [cg.Relation.generate(LINEITEM, per, seed=7, first_row=s * per, nthreads=32) for s in range(a.shards)]
        shards = [cg.Shard(r, [0, 1, 2, 3, 4, 5, 6]) for r in rels]
        run_shapes(lineitem_shapes(), rels, shards, a, want, out, 16)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
