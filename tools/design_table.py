#!/usr/bin/env python
"""Writes the measurement tables of DESIGN.md (section 6: every tracked bench line; section 7: scaling) from
profiles/r02_bench_<cfg>_n<N>.json (one JSON line each, as bench.py printed them), between the
<!-- TABLE:BEGIN --> / <!-- TABLE:END --> and <!-- SCALING:BEGIN --> / <!-- SCALING:END --> markers.
usage: python tools/design_table.py            (rewrites DESIGN.md in place)
       python tools/design_table.py --print    (prints the tables)"""
import glob
import json
import os
import re
import sys

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


def bench_tables(runs):
    out = []
    print = out.append
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
    print("")
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
    return "\n".join(out)


def scaling_table(runs):
    out = ["| cfg | N | ms / step | speed-up over 1 GPU | efficiency | e2e ms (one operator call) | e2e speed-up |", "|---|---|---|---|---|---|---|"]
    for (cfg, n) in sorted(runs):
        if (cfg, 1) not in runs:
            continue
        d, d1 = runs[(cfg, n)], runs[(cfg, 1)]
        sp = d1["ms_per_step"] / d["ms_per_step"]
        e, e1 = d.get("e2e"), d1.get("e2e")
        out.append("| %s | %d | %.3f | %.2f | %.2f | %s | %s |" % (cfg, n, d["ms_per_step"], sp, sp / n,
                   "%.1f" % e["ms_per_step"] if e else "-", "%.2f" % (e1["ms_per_step"] / e["ms_per_step"]) if e and e1 else "-"))
    return "\n".join(out)


def main():
    runs = load()
    tables = {"TABLE": bench_tables(runs), "SCALING": scaling_table(runs)}
    if "--print" in sys.argv:
        print(tables["TABLE"] + "\n\n" + tables["SCALING"])
        return
    path = os.path.join(ROOT, "DESIGN.md")
    text = open(path).read()
    for key, body in tables.items():
        text, n = re.subn(r"<!-- %s:BEGIN -->.*?<!-- %s:END -->" % (key, key),
                          lambda m: "<!-- %s:BEGIN -->\n%s\n<!-- %s:END -->" % (key, body, key), text, flags=re.S)
        if n != 1:
            raise SystemExit("marker pair for %s not found in DESIGN.md" % key)
    open(path, "w").write(text)


if __name__ == "__main__":
    main()
