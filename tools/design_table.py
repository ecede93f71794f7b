#!/usr/bin/env python
"""Prints the measurement tables of DESIGN.md section 6 from the tracked bench lines
profiles/r02_bench_<cfg>_n<N>.json (one JSON line each, as bench.py printed them).
usage: python tools/design_table.py > /tmp/table.md"""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    runs = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_cfg*_n*.json"))):
        m = re.search(r"r02_bench_(cfg\d)_n(\d+)(?:_(\w+))?\.json$", f)
        if not m or m.group(3):
            continue
        try:
            runs[(m.group(1), int(m.group(2)))] = json.loads(open(f).read().strip().splitlines()[-1])
        except Exception:
            pass
    return runs


def main():
    runs = load()
    print("| cfg | N | ms / step (device) | Gpix*frames/s | step: fraction of N x HBM peak | dominant kernel: ms, GB/s, fraction of peak | "
          "e2e ms (operator call) | e2e Gpix*frames/s | CPU reference (cores) | file |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for (cfg, n) in sorted(runs):
        d = runs[(cfg, n)]
        r, e, c = d["roofline"], d.get("e2e"), d.get("cpu_baseline")
        print("| %s | %d | %.3f | %.1f | %.3f | %.3f ms, %.0f GB/s, **%.3f** | %s | %s | %s | `profiles/r02_bench_%s_n%d.json` |" % (
            cfg, n, d["ms_per_step"], d["value"], d["step_hbm_frac"], r["ms_per_launch"], r["achieved"], r["frac"],
            "%.1f" % e["ms_per_step"] if e else "-", "%.1f" % e["value"] if e else "-",
            "%.3f Gpix*f/s (%d)" % (c["value"], c["cores"]) if c and "value" in c else "-", cfg, n))
    print()
    print("| cfg | N | step breakdown (ms) | rejection counters (low, high per layer) | guard | parity of the run |")
    print("|---|---|---|---|---|---|")
    for (cfg, n) in sorted(runs):
        d = runs[(cfg, n)]
        b = ", ".join("%s %.3f" % (k, v) for k, v in d["step_breakdown_ms"].items() if not k.endswith("gbs"))
        p = d.get("parity") or {}
        ptxt = "bit-exact" if p.get("bit_exact") or p.get("bit_exact_vs_oracle_same_sigma_form") else str(p)
        if "max_abs_diff_vs_reference_form" in p:
            ptxt += "; vs reference form: max |d| %d, %.4f %% identical" % (p["max_abs_diff_vs_reference_form"], 100 * p["frac_identical_vs_reference_form"])
        print("| %s | %d | %s | %s | %s | %s |" % (cfg, n, b, d["config"]["rejection_counters"], d["config"]["guard_fired"], ptxt))


if __name__ == "__main__":
    main()
