#!/usr/bin/env python
"""Compact text summary of one `ncu --set full` capture (one kernel launch): the counters profiles/README.md quotes.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [algorithmic_bytes] > profiles/x_full.txt
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy % of peak"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("lts__t_requests_srcunit_tex_op_read.sum", "L2 read requests from SMs"),
    ("lts__t_requests_srcunit_tex_op_red.sum", "L2 reduction requests from SMs"),
    ("lts__t_requests_srcunit_tex_op_write.sum", "L2 write requests from SMs"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps per issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def main():
    rep = sys.argv[1]
    algo = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda k: (vals[hdr.index(k)], units[hdr.index(k)]) if k in hdr else None
    print(f"# {rep.split('/')[-1]}: ncu --set full --clock-control none, one launch")
    print(f"kernel: {vals[hdr.index('Kernel Name')]}")
    dur_us = None
    traffic = 0.0
    for key, label in WANT:
        g = get(key)
        if g is None:
            continue
        v, u = g
        print(f"{label:55s} {v} {u}")
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        if key == "gpu__time_duration.sum":
            dur_us = x * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        if key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            traffic += x * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if dur_us:
        print(f"{'DRAM traffic (read + written)':55s} {traffic / 1e6:.2f} MB = {traffic / dur_us / 1e3:.1f} GB/s")
        if algo:
            print(f"{'algorithmic bytes of the launch':55s} {algo / 1e6:.2f} MB = {algo / dur_us / 1e3:.1f} GB/s; traffic / algorithmic = {traffic / algo:.3f}")


if __name__ == "__main__":
    main()
