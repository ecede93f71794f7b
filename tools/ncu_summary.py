#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, without a GPU) into the text kept under profiles/:
key raw metrics of every captured kernel plus, per kernel, where instructions and stall
samples go (by opcode, and for the stacking kernel by phase: gather+normalize / sort / rejection).
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/NAME.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    print("# ncu summary of", rep)
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        print("\n## kernel:", name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-70s %s %s" % (k, row[i], units[i]))
    src = ncu(rep, "source")
    # the source page concatenates kernels: split on the "Kernel Name" marker rows
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[], hdr=None)
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    for b in blocks:
        hdr, data = b["hdr"], b["rows"]
        iS, iE, iM = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[iE]) for r in data) or 1
        totS = sum(int(r[iM]) for r in data) or 1
        print("\n## per-instruction view:", b["name"])
        print("  SASS instructions: %d   warp-instructions executed: %d   stall samples: %d" % (len(data), tot, totS))
        ops, opsS = collections.Counter(), collections.Counter()
        first = last = None
        for k, r in enumerate(data):
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
            o = (m.group(2) if m else "?").split(".")[0]
            ops[o] += int(r[iE])
            opsS[o] += int(r[iM])
            if "VIMNMX.U16x2" in r[iS]:
                first = k if first is None else first
                last = k
        print("  by opcode (% of executed instructions / % of stall samples):")
        print("   ", ", ".join("%s %.1f/%.1f" % (k, 100 * v / tot, 100 * opsS[k] / totS) for k, v in ops.most_common(18)))
        regions = {"whole kernel": (0, len(data))}
        if first is not None and last - first > 100:
            regions = {"gather+normalize+sums (before the network)": (0, first),
                       "sorting network": (first, last + 1),
                       "rejection + mean + store (after the network)": (last + 1, len(data))}
        for name, (a, z) in regions.items():
            e = sum(int(r[iE]) for r in data[a:z])
            s = sum(int(r[iM]) for r in data[a:z])
            c = collections.Counter()
            for r in data[a:z]:
                for i in stall:
                    c[hdr[i]] += int(r[i] or 0)
            t = sum(c.values()) or 1
            print("  %-48s %5.1f%% of instructions, %5.1f%% of samples; stalls: %s" % (
                name, 100 * e / tot, 100 * s / totS, ", ".join("%s %.0f%%" % (k[6:], 100 * v / t) for k, v in c.most_common(5))))


if __name__ == "__main__":
    main()
