#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200, through libsiril_stack_cuda's C ABI.

Workload (BASELINE.json configs[2], the one `metric` is quoted on; fits one GPU):
    sigma-clip mean (low 4.0 / high 3.0), multiplicative normalization,
    64 synthetic 6000x4000 uint16 RGB frames  (4.608 Gpix*frames, 9.2 GB in, 144 MB out)

One "step" = the whole hot path over the 64 frames:
    K4 histogram of layer 0 of every frame -> [all-reduce(SUM) of the histograms over ranks]
    -> K5 IKSS location/scale -> K6 coefficients -> K3 normalize + sigma-clip + mean
    [-> all-reduce(SUM) of the 6 rejection counters].
`value`     : Gpix*frames/s with the frames resident in HBM (CUDA events on the launch stream).
`e2e`       : the same metric through the operator call a Siril host makes
              (ssc_stack_mean_with_rejection, pinned host frames in, host result out); at N>1 the
              same pipeline per rank on its row band (pinned H2D + kernels + D2H inside the timer).
`roofline`  : the dominant kernel (K3) against the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the reference's own stacking.c + statistics.c (oracle/_ref,
              compiled unmodified) on the host cores, on a bounded row band of the same workload.
Multi-GPU: one rank per GPU (torchrun), the image is split in row bands (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

# DRAM bytes of one K3 launch over the whole benchmark image, measured once under ncu --set full
# (9.253 GB read + 0.164 GB written; the algorithmic bytes are 9.360 GB)
K3_DRAM_BYTES_N1 = 9_252_995_000 + 163_980_288


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n=64, c=3, h=4000, w=6000, rejection=2, sig=(4.0, 3.0), normalize=2, kind=0,
           name="cfg3: sigma-clip mean (4.0/3.0), multiplicative normalization, 64 x 6000x4000 uint16 RGB")
SEED = 0x5349524C
METRIC = "Gpix*frames/s stacked (sigma-clip mean)"
UNIT = "Gpix*frames/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if sm:
            busy = sorted(sm)[len(sm) // 2:]            # the upper half: samples taken under load
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def band_rows(h, world, rank):
    base, rem = divmod(h, world)
    rows = base + (1 if rank < rem else 0)
    row0 = rank * base + min(rank, rem)
    return row0, rows


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own code on the host cores
# ----------------------------------------------------------------------------------------------
def make_sample(rows, pkg=None, torch=None):
    """First `rows` rows of all frames of the workload, as numpy [N, C, rows, W]."""
    n, c, w, h = CFG["n"], CFG["c"], CFG["w"], CFG["h"]
    if torch is not None and torch.cuda.is_available():
        buf = torch.empty((n, c, rows, w), dtype=torch.int16, device="cuda")     # u16 payload
        pkg.dev.synth_frames(buf.data_ptr(), n, c, rows, w, 0, 0, h, CFG["kind"], SEED,
                             torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return buf.cpu().numpy().view(np.uint16)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth                                            # no GPU here: same distribution from numpy
    return synth.make_frames(n, c, rows, w, seed=1, outlier_p=1e-3, zero_p=1e-4, stars=max(1, rows * w // 9216))


def cpu_run(sample):
    """One pass of the reference (compute_normalization + stack_mean_with_rejection) over `sample`."""
    from oracle import refrun
    kind = "reference"
    if refrun.available():
        r = refrun.run(sample, refrun.MEAN, CFG["rejection"], CFG["normalize"], CFG["sig"], want_coeff=True)
        assert r["retval"] == 0
        secs, cores = r["t_method"] + r["t_norm"], refrun.num_procs()
    else:                                                   # oracle/_ref not shipped: the C restatement
        from oracle import port
        kind = "port"
        t0 = time.time()
        st = [port.frame_stats_sorted(f) for f in sample]
        coeff = port.norm_coeffs(CFG["normalize"], [s[0] for s in st], [s[1] for s in st])
        port.stack_mean(sample, CFG["rejection"], CFG["sig"], CFG["normalize"], coeff=coeff)
        secs, cores = time.time() - t0, os.cpu_count()
    return secs, cores, kind


def size_sample(target_s, pkg, torch, cap_rows=1000):
    """Pick a row count whose CPU pass lasts about target_s seconds (probe with 8 rows first)."""
    probe = make_sample(8, pkg, torch)
    cpu_run(probe)                                          # warm the threads / page in the library
    secs, _, _ = cpu_run(probe)
    rows = int(max(8, min(cap_rows, 8 * target_s / max(secs, 1e-3))))
    return rows - rows % 8 if rows >= 16 else rows


def reference_arm(args, pkg, torch):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                              # rank 0 alone runs the CPU arm
    rows = size_sample(3.0, pkg, torch)
    sample = make_sample(rows, pkg, torch)
    units = CFG["n"] * CFG["c"] * rows * CFG["w"] / 1e9
    for _ in range(args.warmup):
        cpu_run(sample)
    t = []
    for _ in range(args.steps):
        secs, cores, kind = cpu_run(sample)
        t.append(secs)
    total = float(np.sum(t))
    value = units * args.steps / total
    desc = "rows 0..%d of all 64 RGB frames (%.3f Gpix*frames per step, normalization statistics on those rows)" % (rows, units)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="u16 samples, f64 (+x87 long double) statistics", data="synthetic", impl="reference",
                config=dict(workload=CFG["name"], frames=CFG["n"], width=CFG["w"], height=CFG["h"], channels=CFG["c"],
                            sample=desc),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=desc),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    _emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def _emit(line):
    """The ONE JSON line, on the real stdout (see main: everything else on fd 1 went to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


_REAL_STDOUT = None


def main():
    # Libraries write to stdout behind our back (NCCL prints its version line there, the reference's
    # stackers printf their block layout): keep fd 1 for the JSON line alone, send the rest to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the end-to-end leg (default min(steps, 5))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--small", action="store_true", help="1/16 of the frame area (debug only; not a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.small:
        CFG.update(h=1000, w=1504, name=CFG["name"] + " [SMALL DEBUG SHAPE]")

    import torch
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    pkg = ge.load_package()

    if args.impl == "reference":
        return reference_arm(args, pkg, torch)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or pkg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    n, c, h, w = CFG["n"], CFG["c"], CFG["h"], CFG["w"]
    row0, rows = band_rows(h, world, rank)
    samples_total = n * c * h * w

    # ---- this rank's row band of every frame, generated in place (counter-based: any band alone) ----
    # (uint16 payloads live in int16 tensors: torch's uint16 kernels are incomplete; the library only sees pointers)
    frames = torch.empty((n, c, rows, w), dtype=torch.int16, device=dev)
    pkg.dev.synth_frames(frames.data_ptr(), n, c, rows, w, 0, row0, h, CFG["kind"], SEED, sp)
    out = torch.empty((c, rows, w), dtype=torch.int16, device=dev)
    hist = torch.zeros((n, pkg.HIST_STRIDE), dtype=torch.int32, device=dev)
    scratch = torch.empty((n, 65536), dtype=torch.int32, device=dev)
    stats = torch.empty((4, n), dtype=torch.float64, device=dev)
    shard_stats = world > 1 and n % world == 0              # reduce-scatter the histograms, IKSS on n / world frames per rank
    n_loc = n // world if shard_stats else n
    if shard_stats:
        hist_loc = torch.empty((n_loc, pkg.HIST_STRIDE), dtype=torch.int32, device=dev)
        stats_loc = torch.empty((4, n_loc), dtype=torch.float64, device=dev)
        stats_all = torch.empty((world, 4, n_loc), dtype=torch.float64, device=dev)
    coeff = torch.empty((3, n), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    counters = torch.zeros(pkg.N_COUNTERS, dtype=torch.int64, device=dev)
    band = pkg.dev.band(frames.data_ptr(), n, c, h, w, out_row0=row0, out_rows=rows, buf_row0=row0, buf_rows=rows)
    launches_per_step = 6 if c > 1 else 5                   # K4 (+ other-layer check), K5, K6, K3, K3 hand-backs

    def step(ev=None):
        hist.zero_()
        counters.zero_()
        if ev: ev[0].record(stream)
        pkg.dev.histogram(band, hist.data_ptr(), sp)
        if world > 1 and shard_stats:
            # exact integer SUM over the row bands, each rank receiving the summed histograms of ITS frames
            # only (half the traffic of an all-reduce); it runs the IKSS for those, then the per-frame
            # statistics (4 doubles each) are gathered: every rank ends up with identical coefficients
            dist.reduce_scatter_tensor(hist_loc.view(-1), hist.view(-1))
            if ev: ev[1].record(stream)
            pkg.dev.ikss(hist_loc.data_ptr(), n_loc, stats_loc.data_ptr(), scratch.data_ptr(), sp)
            dist.all_gather_into_tensor(stats_all.view(-1), stats_loc.view(-1))
            stats.copy_(stats_all.view(world, 4, n_loc).permute(1, 0, 2).reshape(4, n))
        else:
            if world > 1:
                dist.all_reduce(hist)                       # exact integer SUM over the row bands (NCCL)
            if ev: ev[1].record(stream)
            pkg.dev.ikss(hist.data_ptr(), n, stats.data_ptr(), scratch.data_ptr(), sp)
        pkg.dev.norm_coeffs(stats.data_ptr(), n, 0, CFG["normalize"], coeff.data_ptr(), status.data_ptr(), sp)
        if ev: ev[2].record(stream)
        pkg.dev.stack_mean(band, out.data_ptr(), counters.data_ptr(), CFG["rejection"], CFG["sig"],
                           CFG["normalize"], coeff.data_ptr(), stream=sp)
        if ev: ev[3].record(stream)
        if world > 1:
            dist.all_reduce(counters)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps ----
    for _ in range(args.warmup):
        step()
    sync_all()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for k in range(args.steps):
        step(evs[k])
    t1.record(stream)
    sync_all()
    ms_total = t0.elapsed_time(t1)
    phase = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in evs])   # hist(+allreduce), ikss+coeff, stack
    tt = torch.tensor([ms_total] + phase.mean(0).tolist(), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)           # max over ranks
    ms_total, ms_hist, ms_ikss, ms_stack = tt.tolist()
    ms_step = ms_total / args.steps
    value = samples_total / 1e9 / (ms_step / 1e3)
    assert int(status.item()) == 0
    irej = counters[:6].cpu().numpy().astype(np.uint64).reshape(3, 2)
    guard, cap = int(counters[6].item()), int(counters[7].item())

    # ---- end to end: pinned host frames -> result on the host, copies inside the timer ----
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 5)
    host = torch.empty((n, c, rows, w), dtype=torch.int16, pin_memory=True)
    host.copy_(frames)
    torch.cuda.synchronize()
    h2d_bytes = host.numel() * 2
    d2h_bytes = c * rows * w * 2
    if world == 1:
        seq = pkg.Sequence([host[i].numpy().view(np.uint16) for i in range(n)])

        def e2e_step():
            r = pkg.stack_mean_with_rejection(seq, CFG["rejection"], CFG["sig"], CFG["normalize"], want_coeff=False)
            return r
        e2e_api = "ssc_stack_mean_with_rejection (operator layer, whole image)"
    else:
        host_out = torch.empty((c, rows, w), dtype=torch.int16, pin_memory=True)

        def e2e_step():
            frames.copy_(host, non_blocking=True)
            step()
            host_out.copy_(out, non_blocking=True)
            return None
        e2e_api = "per-rank band pipeline: pinned H2D + ssc_dev_* + D2H"
    for _ in range(2):
        r = e2e_step()
    sync_all()
    t_wall0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = e2e_step()
    sync_all()
    e2e_ms = (time.perf_counter() - t_wall0) * 1e3 / max(e2e_steps, 1)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_value = samples_total / 1e9 / (e2e_ms / 1e3)
    e2e_parity = None
    if world == 1 and r is not None:
        e2e_parity = bool(np.array_equal(r.data, out.cpu().numpy().view(np.uint16))) and bool(np.array_equal(r.irej, irej))
        op_timing = dict(h2d_ms=r.t_h2d_ms, device_ms=r.t_device_ms, d2h_ms=r.t_d2h_ms, total_ms=r.t_total_ms)
    else:
        op_timing = None

    clocks = sampler.stop() if sampler else None

    # ---- parity spot check of this very run against the oracle (not timed, small sample) ----
    parity = None
    if rank == 0:
        try:
            from oracle import port
            ys = np.linspace(0, rows - 1, 6).astype(int)
            yt = torch.as_tensor(ys, device=dev)
            sub = frames.index_select(2, yt).cpu().numpy().view(np.uint16)    # [N, C, 6, W]
            cf = coeff.cpu().numpy()
            chk = port.stack_mean(sub, CFG["rejection"], CFG["sig"], CFG["normalize"], coeff=cf, sd_mode=port.SD_EXACT)
            got = out.index_select(1, yt).cpu().numpy().view(np.uint16)
            gsl = port.stack_mean(sub, CFG["rejection"], CFG["sig"], CFG["normalize"], coeff=cf, sd_mode=port.SD_GSL)
            d = np.abs(got.astype(np.int32) - gsl["out"].astype(np.int32))
            parity = dict(rows_checked=len(ys), bit_exact_vs_oracle_same_sigma_form=bool(np.array_equal(got, chk["out"])),
                          max_abs_diff_vs_reference_form=int(d.max()), frac_identical_vs_reference_form=float((d == 0).mean()))
        except Exception as e:                                           # the checker is optional at bench time
            parity = dict(error=str(e))

    # ---- CPU baseline: the reference's own code on this box's cores, bounded sample (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            crow = size_sample(12.0, pkg, torch)
            sample = make_sample(crow, pkg, torch)
            secs, cores, kind = cpu_run(sample)
            units = n * c * crow * w / 1e9
            cpu = dict(value=units / secs, unit=UNIT, cores=cores, kind=kind, seconds=secs,
                       sample="rows 0..%d of all 64 RGB frames (%.3f Gpix*frames, statistics on those rows), one pass" % (crow, units))
        except Exception as e:
            cpu = dict(error=str(e))

    if rank == 0:
        peak, peak_src = measured_peaks()
        rows_max = band_rows(h, world, 0)[1]
        bytes_stack = 2 * n * c * rows_max * w + 2 * c * rows_max * w    # per launch, on the largest band
        bytes_hist = 2 * n * rows_max * w
        achieved = bytes_stack / (ms_stack / 1e3) / 1e9
        bytes_alg_step = (2 * n * c * h * w + 2 * c * h * w + 2 * n * h * w)
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_step, higher_is_better=True, scaling="strong", vs_baseline=None,
            dtype="u16 samples, f64 clipping/normalization arithmetic", data="synthetic",
            config=dict(workload=CFG["name"], frames=n, width=w, height=h, channels=c,
                        parallelism="row bands x%d" % world, l2="inputs (9.2 GB) exceed L2 (126 MB): no flush needed",
                        seed=hex(SEED), rejection_counters=irej.tolist(), guard_fired=guard, cap_hit=cap),
            roofline=dict(bound="hbm", kernel="k_stack_fast<64,true,SIGMA_REG> (K3: normalize + sigma-clip + mean, register sorting network)",
                          achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=K3_DRAM_BYTES_N1 // world,
                          traffic_source="dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the "
                                         "whole-image launch (profiles/r01_ncu_k_stack_fast_full_size_launch.txt), / ranks",
                          peak_source=peak_src, algorithmic_bytes_per_launch=bytes_stack,
                          ms_per_launch=ms_stack),
            step_breakdown_ms=dict(histogram_allreduce=ms_hist, ikss_coeffs=ms_ikss, stack=ms_stack,
                                   histogram_gbs=bytes_hist / (ms_hist / 1e3) / 1e9),
            step_hbm_frac=bytes_alg_step / world / (ms_step / 1e3) / 1e9 / peak,
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d_bytes * world, d2h_bytes_per_step=d2h_bytes * world,
                     ms_per_step=e2e_ms, steps=e2e_steps, api=e2e_api, operator_timing_ms=op_timing,
                     matches_device_path=e2e_parity),
            gpu_launches=launches_per_step * args.steps,
            clocks=clocks, parity=parity, cpu_baseline=cpu)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
