#!/usr/bin/env python
"""A few steps of the benchmark's hot path on a reduced number of rows, for ncu captures
(ncu replays every kernel ~40 times: keep it short).  Same kernels, same shapes per row as
bench.py's workload (64 RGB frames, 6000 px wide, sigma clip + multiplicative normalization)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--width", type=int, default=6000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rejection", type=int, default=2)
    ap.add_argument("--normalize", type=int, default=2)
    ap.add_argument("--median", action="store_true")
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="ssc_set_kernel_path: 0 auto, 1 general, 2 pipeline kernel")
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if a.generic:
        pkg.set_kernel_path(1)
    elif a.path:
        pkg.set_kernel_path(a.path)
    n, c, h, w = a.frames, a.layers, a.rows, a.width
    dev = torch.device("cuda", 0)
    sp = torch.cuda.current_stream().cuda_stream
    frames = torch.empty((n, c, h, w), dtype=torch.int16, device=dev)
    pkg.dev.synth_frames(frames.data_ptr(), n, c, h, w, 0, 0, h, 0, 0x5349524C, sp)
    out = torch.empty((c, h, w), dtype=torch.int16, device=dev)
    hist = torch.zeros((n, pkg.HIST_STRIDE), dtype=torch.int32, device=dev)
    scratch = torch.empty((n, 65536), dtype=torch.int32, device=dev)
    stats = torch.empty((4, n), dtype=torch.float64, device=dev)
    coeff = torch.empty((3, n), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    counters = torch.zeros(pkg.N_COUNTERS, dtype=torch.int64, device=dev)
    band = pkg.dev.band(frames.data_ptr(), n, c, h, w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(a.steps):
        hist.zero_()
        counters.zero_()
        e0.record()
        if a.normalize:
            pkg.dev.histogram(band, hist.data_ptr(), sp)
            pkg.dev.ikss(hist.data_ptr(), n, stats.data_ptr(), scratch.data_ptr(), sp)
            pkg.dev.norm_coeffs(stats.data_ptr(), n, 0, a.normalize, coeff.data_ptr(), status.data_ptr(), sp)
        if a.median:
            pkg.dev.stack_median(band, out.data_ptr(), a.normalize, coeff.data_ptr(), stream=sp)
        else:
            pkg.dev.stack_mean(band, out.data_ptr(), counters.data_ptr(), a.rejection, (4.0, 3.0),
                               a.normalize, coeff.data_ptr(), stream=sp)
        e1.record()
        torch.cuda.synchronize()
        print("step %d: %.3f ms  counters %s" % (k, e0.elapsed_time(e1), counters.tolist()))


if __name__ == "__main__":
    main()
