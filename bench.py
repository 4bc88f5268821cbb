#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1] / [2], SURVEY.md 8(d) "C2/C3"): a 1 B-row table of 8 int8
columns in 32 hash-distributed columnar shards (stripe 150 000, chunk group 10 000,
compression none, no NULLs), query
    SELECT key, sum(v), count(*) FROM t WHERE f < 50 GROUP BY key          -- 1 M groups
One step = one execution of that query over all shards: per-shard fused scan+filter+partial
aggregate on the GPU that owns the shard (shard s -> rank s mod N), then the coordinator
combine (NCCL reduce over NVLink at N > 1) and the compaction of the result rows.

  value   rows/s with the shards resident in HBM when the timed region starts
  e2e     rows/s through the C-ABI call on HOST page images (cg_scan_relation): pinned staging
          + cudaMemcpyAsync + kernel + device->host fetch of the result, all timed
  --impl reference  the CPU restatement of the reference's path (oracle/, one thread per
          shard) on the box's host cores -- NOT PostgreSQL/Citus, which is not installed here
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260922
NSHARDS = 32
STRIPE_ROWS = 150_000
CHUNK_ROWS = 10_000
NKEYS = 1_000_000
COLUMNS = [(8, 0, 0, NKEYS, 0), (8, 0, 0, 100, 0), (8, 0, -10**9, 10**9, 0)] + [(8, 0, 0, 1 << 40, 0)] * 5
QUALS = [(1, "<", 50)]
GROUP = [0]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=1_000_000_000)
    p.add_argument("--resident", default="all", choices=["all", "projected"],
                   help="stage all 8 columns in HBM (default) or only the 3 the query reads")
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--pageable", action="store_true", help="e2e from pageable host pages (host de-framing) instead of pinned pages (DMA)")
    p.add_argument("--gen-threads", type=int, default=0)
    p.add_argument("--compression", default="none", choices=["none", "lz4", "zstd"],
                   help="columnar.compression of the synthetic shards (BASELINE configs use none; lz4 exercises the GPU decoder)")
    p.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPUs of its GPU's NUMA node")
    return p.parse_args()


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU during the timed region (NVML)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:          # noqa
            log("clock sampling unavailable:", e)

    def sample_once(self):
        """one NVML sample from the calling thread (used while the GPU is still executing the timed steps)"""
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                              (0x4, "sw_power_cap"), (0x80, "hw_power_brake")):
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        self.stop_flag.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def generate_shards(cg, shard_ids, rows_per_shard, nthreads):
    rels = {}
    t0 = time.time()
    for s in shard_ids:
        rels[s] = cg.Relation.generate(COLUMNS, rows_per_shard, seed=SEED, first_row=s * rows_per_shard,
                                       stripe_row_limit=STRIPE_ROWS, chunk_row_limit=CHUNK_ROWS, nthreads=nthreads)
    log(f"generated {len(shard_ids)} shards x {rows_per_shard} rows in {time.time() - t0:.1f}s")
    return rels


# ----------------------------------------------------------------------------- CPU arm
def cpu_scan(rels, rows_per_shard, nthreads, budget_s=20.0):
    """the oracle's restatement of the reference path, one thread per shard (the reference fans
    one task per shard out over <= citus.max_adaptive_executor_pool_size connections).  Returns
    (rows processed, seconds, per-shard oracle results, sample description)."""
    from oracle import oracle as orc
    orc.build()
    aggs = [orc.sum_(2), orc.count_star()]
    shard_ids = sorted(rels)
    attlen = [c[0] for c in COLUMNS]

    def attach(rel, nstripes):
        v = rel.view
        return orc.Table.attach_view(v.pages, v.nblocks, C.cast(v.stripes, C.c_void_p), nstripes,
                                     C.cast(v.nodes, C.c_void_p), v.nnodes, attlen, chunk_row_limit=CHUNK_ROWS)

    # calibrate on a few stripes of one shard
    total_stripes = rels[shard_ids[0]].view.nstripes
    probe = min(4, total_stripes)
    t0 = time.time()
    attach(rels[shard_ids[0]], probe).scan(QUALS, GROUP, aggs)
    per_stripe = (time.time() - t0) / probe
    waves = -(-len(shard_ids) // nthreads)
    use = total_stripes
    if per_stripe * total_stripes * waves > budget_s:
        use = max(1, int(budget_s / (per_stripe * waves)))
    results = {}

    def work(s):
        results[s] = attach(rels[s], use).scan(QUALS, GROUP, aggs)

    t0 = time.time()
    pending = list(shard_ids)
    while pending:
        batch, pending = pending[:nthreads], pending[nthreads:]
        ths = [threading.Thread(target=work, args=(s,)) for s in batch]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    secs = time.time() - t0
    rows = sum(r.rows_scanned for r in results.values())
    sample = (f"{len(shard_ids)} shards x {use}/{total_stripes} stripes = {rows} rows, one thread per shard "
              f"({min(nthreads, len(shard_ids))} threads)")
    return rows, secs, results, sample


def run_reference(args, rank, world):
    if rank != 0:
        return
    from citus_b200 import build
    build.build()
    from citus_b200 import columnar as cg     # host-side generator only (no GPU call)
    rows_per_shard = args.rows // NSHARDS
    ncpu = os.cpu_count() or 8
    cg.set_writer_compression(args.compression)
    rels = generate_shards(cg, range(NSHARDS), rows_per_shard, args.gen_threads or min(64, ncpu))
    cg.set_writer_compression("none")
    nthreads = min(NSHARDS, max(1, ncpu // 2))
    budget = 12.0
    times = []
    rows = 0
    sample = ""
    for i in range(args.warmup + args.steps):
        rows, secs, _, sample = cpu_scan(rels, rows_per_shard, nthreads, budget_s=budget)
        if i >= args.warmup:
            times.append(secs)
        if i == 0 and args.warmup + args.steps > 6:
            budget = 6.0
    secs = float(np.mean(times))
    value = rows / secs
    line = {
        "impl": "reference", "metric": "rows/sec scan+GROUP BY over 1B-row columnar shards", "value": value,
        "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, f"shard s -> GPU s mod {args.gpus}"),
        "reference_arm": "host cores, one thread per shard",
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference path (oracle/oracle.c): PostgreSQL/Citus cannot be built here",
    }
    emit(line)


def workload_config(args, where):
    return {"workload": "C2: 32 columnar shards, 1B rows x 8 int8 cols, WHERE f<50 GROUP BY key (1M keys), sum(v), count(*)",
            "rows": args.rows, "shards": NSHARDS, "stripe_row_limit": STRIPE_ROWS, "chunk_group_row_limit": CHUNK_ROWS,
            "groups": NKEYS, "selectivity": 0.5, "compression": args.compression, "parallelism": where,
            "l2_policy": "inputs (24.4 GB/step) are far larger than the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from citus_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from citus_b200 import capi, columnar as cg
    from citus_b200 import distributed as cgd
    torch.cuda.set_device(local_rank)
    cg.init(local_rank)
    numa_node = -1 if args.no_numa_bind else cg.numa_bind()    # page images are first-touched on the GPU's socket
    # one non-default torch stream for everything: the library's kernels, torch ops, the events that
    # time them and (through torch's stream dependencies) the NCCL collectives
    torch.cuda.set_stream(torch.cuda.Stream())
    cg.use_torch_stream()

    rows_per_shard = args.rows // NSHARDS
    total_rows = rows_per_shard * NSHARDS
    my_shards = cgd.shards_of_rank(NSHARDS, rank, world)
    ncpu = os.cpu_count() or 8
    gen_threads = args.gen_threads or max(4, min(64, ncpu // max(world, 1)))
    cg.set_writer_compression(args.compression)
    rels = generate_shards(cg, my_shards, rows_per_shard, gen_threads)
    cg.set_writer_compression("none")

    aggs = [cg.sum_(2), cg.count_star()]
    desc = cg.make_desc(QUALS, GROUP, aggs)
    kmin, kmax, bound = None, None, 0
    for rel in rels.values():
        a, b, bounds, _ = cg.relation_bounds(rel, desc)
        kmin = a if kmin is None else min(kmin, a)
        kmax = b if kmax is None else max(kmax, b)
        bound = max(bound, bounds[0])
    if world > 1:                      # identical table layout on every rank
        t = torch.tensor([-kmin, kmax, bound], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kmin, kmax, bound = -int(t[0]), int(t[1]), int(t[2])
    aggs[0].term_abs_bound = bound
    desc = cg.make_desc(QUALS, GROUP, aggs)
    coldescs = rels[my_shards[0]].column_descs()
    partial = cg.GpuColumnarAgg(desc, coldescs, kmin, kmax, total_rows)
    nw, ops, dense, cap = partial.layout()
    log(f"rank {rank}: group table dense={dense} capacity={cap} words={nw}")

    t0 = time.time()
    stage_cols = None if args.resident == "all" else [0, 1, 2]
    shards = {s: cg.Shard(rels[s], stage_cols) for s in my_shards}
    resident = sum(sh.device_bytes for sh in shards.values())
    log(f"rank {rank}: staged {len(shards)} shards, {resident / 1e9:.1f} GB resident in {time.time() - t0:.1f}s")

    out_keys = torch.empty(NKEYS + 2, dtype=torch.int64, device="cuda")
    out_nulls = torch.empty(NKEYS + 2, dtype=torch.uint8, device="cuda")
    out_words = torch.empty((NKEYS + 2) * nw, dtype=torch.int64, device="cuda")

    def step():
        try:
            partial.reset()
            for s in my_shards:
                partial.scan_shard(shards[s], want_stats=False)
            cgd.combine_partials(partial, dst=0)
            if rank == 0:
                return partial.export_device(out_keys.data_ptr(), out_nulls.data_ptr(), out_words.data_ptr(), NKEYS + 2)
            return 0
        except capi.CitusGpuError as e:
            if e.code != capi.CG_ERETRY_UNPACKED or world > 1:
                raise
            log("packed accumulators overflowed: falling back to the two-word path")
            partial.reset()
            partial.set_packing(False)
            return step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # algorithmic bytes per launch and a first parity look (one instrumented step)
    partial.reset()
    algo_bytes = []
    rows_scanned = 0
    for s in my_shards:
        st = partial.scan_shard(shards[s], want_stats=True)
        algo_bytes.append(st.bytes_scanned)
        rows_scanned += st.rows_scanned
    assert rows_scanned == rows_per_shard * len(my_shards)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    capi.check(capi.lib().cg_profile_begin())
    launches_before = capi.lib().cg_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    ngroups = 0
    for _ in range(args.steps):
        ngroups = step()
    ev1.record()
    sampler.sample_once()                      # the steps are short: sample while the device is still inside them
    while not ev1.query():
        sampler.sample_once()
    barrier()
    clocks = sampler.result()
    all_launches = capi.lib().cg_kernel_launches() - launches_before
    launches, ktotal, kmax_ms = C.c_int32(), C.c_double(), C.c_double()
    capi.check(capi.lib().cg_profile_collect(C.byref(launches), C.byref(ktotal), C.byref(kmax_ms)))
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    value = total_rows / (ms / 1e3)

    peak, peak_src = measured_peak()
    avg_kernel_ms = ktotal.value / max(launches.value, 1)
    avg_bytes = float(np.mean(algo_bytes))
    achieved = avg_bytes / (avg_kernel_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "cg_scan_fast_kernel<1,DENSE,1> (fused decode+filter+partial aggregate)",
                "bytes_per_launch": avg_bytes, "avg_launch_ms": avg_kernel_ms, "peak_source": peak_src,
                "kernel_share_of_step": ktotal.value / args.steps / ms}
    # ncu DRAM traffic per launch, when a capture summary has been committed
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- end to end on host buffers
    e2e = None
    if not args.no_e2e:
        h2d = 0
        t0 = time.time()
        if not args.pageable:
            for s in my_shards:               # pin the page images once (what registering shared_buffers would do)
                rels[s].register()
        reg_s = time.time() - t0
        log(f"rank {rank}: registered {len(my_shards)} page images in {reg_s:.1f}s")
        partial.reset()
        for s in my_shards:                   # instrumented warm-up pass: bytes moved
            h2d += partial.scan_relation(rels[s], want_stats=True).h2d_bytes
        d2h = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.e2e_steps):
            partial.reset()
            for s in my_shards:
                partial.scan_relation(rels[s], want_stats=False)
            cgd.combine_partials(partial, dst=0)
            if rank == 0:
                res = partial.fetch(reuse=True)   # device -> host read of the result rows, into caller-owned buffers
                d2h = res["n"] * (9 + 8 * nw)
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1) / args.e2e_steps
        if world > 1:
            t = torch.tensor([ems, float(h2d)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
            ems, h2d = float(t[0]), int(t[1])
        e2e = {"value": total_rows / (ems / 1e3), "unit": "rows/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ems, "steps": args.e2e_steps,
               "host_buffers": "pageable pages -> host de-frame into pinned blocks -> cudaMemcpyAsync" if args.pageable else
                               "pinned (cudaHostRegister) page images -> 1-D DMA of whole pages -> GPU drops page headers + realigns",
               "api": "cg_scan_relation + cg_partial_fetch", "numa_node": numa_node}

    # ---- CPU baseline + full-size parity (rank 0, N = 1 only)
    cpu = None
    parity = None
    cg.numa_unbind()                          # the CPU leg may use every core of the box
    if rank == 0 and world == 1 and not args.no_cpu:
        nthreads = min(NSHARDS, max(1, ncpu // 2))
        rows, secs, results, sample = cpu_scan(rels, rows_per_shard, nthreads)
        cpu = {"value": rows / secs, "unit": "rows/s", "cores": nthreads, "kind": "port", "sample": sample}
        if rows == total_rows:                # the oracle covered the whole workload: compare every group
            from oracle import oracle as orc
            merged = None
            for s in sorted(results):
                if merged is None:
                    merged = results[s]
                else:
                    merged.combine(results[s])
            step()
            got = partial.groups()
            want = merged.groups()
            same = set(got) == set(want) and all(
                got[k][0]["sum"] == want[k][0]["sum"] and got[k][1]["count"] == want[k][1]["count"] for k in want)
            parity = {"groups": len(want), "bit_exact": bool(same)}
            if not same:
                log("PARITY FAILURE at full size")

    if rank == 0:
        line = {
            "metric": "rows/sec scan+GROUP BY over 1B-row columnar shards", "value": value, "unit": "rows/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "config": workload_config(args, f"shard s -> GPU s mod {world}"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(all_launches), "scan_kernel_launches": int(launches.value),
            "roofline": roofline,
            "cpu_baseline": cpu, "groups": int(ngroups), "resident_bytes_rank0": int(resident),
            "hbm_gbs_whole_step": total_rows * 24.375 / (ms / 1e3) / 1e9 / world,
        }
        if parity is not None:
            line["parity_full_size"] = parity
        emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, library
    chatter) was redirected to stderr at start-up"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
