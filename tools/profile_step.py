#!/usr/bin/env python
"""A few steps of one BASELINE workload's hot path on a reduced number of rows, for ncu captures
(ncu replays every kernel ~40 times: keep it short).  Same kernels, same shapes per row as
bench.py --config <cfg> (tools/workload.py)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import workload as wl                                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3", choices=sorted(wl.CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="rows of the band (default: the whole image)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rejection", type=int, default=None)
    ap.add_argument("--normalize", type=int, default=None)
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--median", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="ssc_set_kernel_path: 0 auto, 1 general kernel")
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if a.path:
        pkg.set_kernel_path(a.path)
    cfg = dict(wl.CONFIGS[a.config])
    if a.rejection is not None:
        cfg["rejection"] = a.rejection
        cfg["sig"] = (5.0, 2.5) if a.rejection == 5 else ((0.2, 0.1) if a.rejection == 1 else (4.0, 3.0))
    if a.normalize is not None:
        cfg["normalize"] = a.normalize
    if a.frames:
        cfg["n"] = a.frames
    if a.median:
        cfg["op"] = "median"
    n, c, w, full_h = cfg["n"], cfg["c"], cfg["w"], cfg["h"]
    h = a.rows or full_h
    row0 = (full_h - h) // 2
    dev = torch.device("cuda", 0)
    sp = torch.cuda.current_stream().cuda_stream
    frames = torch.empty((n, c, h, w), dtype=torch.int16, device=dev)
    pkg.dev.synth_frames(frames.data_ptr(), n, c, h, w, 0, row0, full_h, cfg["kind"], wl.SEED, sp)
    out = torch.empty((c, h, w), dtype=torch.int16, device=dev)
    hist = torch.zeros((n, pkg.HIST_STRIDE), dtype=torch.int32, device=dev)
    scratch = torch.empty((n, 65536), dtype=torch.int32, device=dev)
    stats = torch.empty((4, n), dtype=torch.float64, device=dev)
    coeff = torch.empty((3, n), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    counters = torch.zeros(pkg.N_COUNTERS, dtype=torch.int64, device=dev)
    acc = torch.empty((c, h, w), dtype=torch.int32, device=dev)
    maxim = torch.zeros(1, dtype=torch.int64, device=dev)
    band = pkg.dev.band(frames.data_ptr(), n, c, h, w, row_order=cfg["row_order"])
    hband = pkg.dev.band(frames.data_ptr(), n, c, h, w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(a.steps):
        hist.zero_()
        counters.zero_()
        maxim.zero_()
        e0.record()
        if cfg["op"] == "sum":
            pkg.dev.sum_accumulate(band, acc.data_ptr(), maxim.data_ptr(), sp)
            pkg.dev.sum_finish(acc.data_ptr(), c * h * w, maxim.data_ptr(), out.data_ptr(), sp)
        else:
            if cfg["normalize"]:
                pkg.dev.histogram(hband, hist.data_ptr(), sp)
                pkg.dev.ikss(hist.data_ptr(), n, stats.data_ptr(), scratch.data_ptr(), sp)
                pkg.dev.norm_coeffs(stats.data_ptr(), n, 0, cfg["normalize"], coeff.data_ptr(), status.data_ptr(), sp)
            if cfg["op"] == "median":
                pkg.dev.stack_median(band, out.data_ptr(), cfg["normalize"], coeff.data_ptr(), stream=sp)
            else:
                pkg.dev.stack_mean(band, out.data_ptr(), counters.data_ptr(), cfg["rejection"], cfg["sig"],
                                   cfg["normalize"], coeff.data_ptr(), stream=sp)
        e1.record()
        torch.cuda.synchronize()
        print("%s rows %d step %d: %.3f ms  counters %s" % (a.config, h, k, e0.elapsed_time(e1), counters.tolist()))


if __name__ == "__main__":
    main()
