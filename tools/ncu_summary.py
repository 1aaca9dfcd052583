#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the text summary committed under profiles/.

usage: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>.txt
(runs `ncu -i ... --page raw --csv` and `--page source --csv --print-source sass`; no GPU needed)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True, check=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-74s %s %s" % (k, r[i], units[i]))
        st = []
        for i, k in enumerate(hdr):
            if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued"):
                try:
                    st.append((float(r[i].replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("  warp stall samples (all):")
        for v, k in sorted(st, reverse=True)[:10]:
            print("    %-28s %7.0f  %5.1f %%" % (k, v, 100 * v / tot))
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h = src[1]
    isrc, ismp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    byop, exop = collections.Counter(), collections.Counter()
    for r in src[2:]:
        t = r[isrc].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += int(r[ismp] or 0)
        exop[op] += int(r[iex] or 0)
    tot = sum(byop.values()) or 1
    print("  SASS opcode mix (warp-level executions, share of stall samples):")
    for op, n in exop.most_common(14):
        print("    %-10s %12d  %5.1f %%" % (op, n, 100.0 * byop[op] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
