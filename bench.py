#!/usr/bin/env python
"""bench.py -- the reference's stacking workloads on B200, through libsiril_stack_cuda's C ABI.

Default workload (BASELINE.json configs[2], the one `metric` is quoted on; fits one GPU):
    sigma-clip mean (low 4.0 / high 3.0), multiplicative normalization,
    64 synthetic 6000x4000 uint16 RGB frames  (4.608 Gpix*frames, 9.2 GB in, 144 MB out)
`--config cfg1|cfg2|cfg3|cfg4|cfg5` selects any of BASELINE.json's five shapes (tools/workload.py).

One "step" = the whole hot path over all frames of the workload:
    mean, normalized  : K4 histogram of layer 0 -> [histograms summed over ranks] -> K5 IKSS -> K6
                        coefficients -> K3 normalize + reject + mean [-> rejection counters summed]
    mean, unnormalized: K3                      median: K2
    sum               : K1 accumulate -> [max over ranks] -> K1b rescale
`value`     : Gpix*frames/s with the frames resident in HBM (CUDA events on the launch stream).
`e2e`       : the same metric through the operator call a Siril host makes (ssc_stack_*: pinned host
              frames in, host result out); at N>1 the same pipeline per rank on its row band.
`roofline`  : the workload's dominant kernel against the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the reference's own stacking.c + statistics.c (oracle/_ref,
              compiled unmodified) on the host cores, on a bounded row band of the same workload; that
              process never loads the product library.
Multi-GPU: one rank per GPU (torchrun), the image is split in row bands (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tools import workload as wl                                       # noqa: E402

SEED = wl.SEED
UNIT = "Gpix*frames/s"
L2_BYTES = 126 << 20

# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel's launch over the whole
# image on one GPU, from one `ncu --set full` capture each (profiles/): None where not captured
DRAM_BYTES_N1 = {
    "cfg1": (33_564_928, "profiles/r02_ncu_cfg1_sum.txt (k_sum_rows_vec8, whole image)"),
    "cfg2": (1_074_024_000 + 34_832_896, "profiles/r02_ncu_cfg2_median.txt (whole image)"),
    "cfg3": (9_252_995_000 + 163_980_288, "profiles/r01_ncu_k_stack_fast_full_size_launch.txt (whole image)"),
    "cfg4": (8 * (822_657_792 + 65_631_488), "profiles/r02_ncu_cfg4_mid_v2.txt (440-row slice = 1/8 of the image, x 8)"),
    "cfg5": (8 * (964_913_408 + 232_188_160), "profiles/r02_ncu_cfg5_warp_literal_pass.txt (152-row slice = 1/8 of the image, x 8; includes the kernel's local-memory traffic)"),
}
DOMINANT = {
    "cfg1": "k_sum_rows_vec8 (K1: per-pixel u32 sums + max)",
    "cfg2": "k_stack_fast<32,true,MEDIAN> (K2: register sorting network, truncating median)",
    "cfg3": "k_stack_fast<64,true,SIGMA_REG> (K3: normalize + sigma-clip + mean, register sorting network)",
    "cfg4": "k_stack_mid<256,64> (K3: normalize + sort in shared memory + winsorized clipping + mean)",
    "cfg5": "k_stack_warp<32> (K3: one warp per pixel pair, register bitonic sort over the lanes, exact-sums linear-fit clipping)",
}


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


def metric_name(key):
    return "Gpix*frames/s stacked (%s)" % wl.MODE_NAME[key]


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own code on the host cores.  Nothing here touches the product library.
# ----------------------------------------------------------------------------------------------
def cpu_run(cfg, sample):
    """One pass of the reference operator of the workload over `sample` (uint16 [N, C, rows, W])."""
    from oracle import refrun
    if refrun.available():
        method = dict(sum=refrun.SUM, median=refrun.MEDIAN, mean=refrun.MEAN)[cfg["op"]]
        r = refrun.run(sample, method, cfg["rejection"], cfg["normalize"], cfg["sig"], row_order=cfg["row_order"],
                       want_coeff=True)
        assert r["retval"] == 0
        # stack_summing has no OpenMP pragma at all (stacking.c:196-355): one core whatever the machine
        return r["t_method"] + r["t_norm"], (1 if cfg["op"] == "sum" else refrun.num_procs()), "reference"
    from oracle import port                                             # oracle/_ref not shipped: the C restatement
    t0 = time.time()
    if cfg["op"] == "sum":
        port.stack_sum(sample, row_order=cfg["row_order"])
    elif cfg["op"] == "median":
        port.stack_median(sample, row_order=cfg["row_order"])
    else:
        coeff = None
        if cfg["normalize"]:
            st = [port.frame_stats_sorted(f) for f in sample]
            coeff = port.norm_coeffs(cfg["normalize"], [s[0] for s in st], [s[1] for s in st])
        port.stack_mean(sample, cfg["rejection"], cfg["sig"], cfg["normalize"], coeff=coeff, row_order=cfg["row_order"])
    return time.time() - t0, os.cpu_count(), "port"


def cpu_sample_rows(cfg, budget_s, passes):
    """Rows of the image whose `passes` CPU passes last about budget_s seconds in total (probe first)."""
    probe_rows = 8
    probe = wl.host_frames(cfg, cfg["h"] // 2, probe_rows)
    cpu_run(cfg, probe)                                                 # warm the threads / page in the library
    secs, _, _ = cpu_run(cfg, probe)
    rows = int(probe_rows * budget_s / max(passes, 1) / max(secs, 1e-4))
    rows = max(8, min(cfg["h"], rows))
    return rows - rows % 8 if rows < cfg["h"] else rows


def sample_desc(cfg, row0, rows):
    units = cfg["n"] * cfg["c"] * rows * cfg["w"] / 1e9
    whole = rows == cfg["h"]
    return units, ("the whole image, one pass = the full configuration" if whole else
                   "storage rows %d..%d of %d of all %d frames (%.3f Gpix*frames per pass; normalization statistics "
                   "on those rows)" % (row0, row0 + rows, cfg["h"], cfg["n"], units))


def reference_arm(args, key, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                              # rank 0 alone runs the CPU arm
    rows = cpu_sample_rows(cfg, 150.0, args.steps + args.warmup)
    row0 = (cfg["h"] - rows) // 2
    sample = wl.host_frames(cfg, row0, rows)
    units, desc = sample_desc(cfg, row0, rows)
    for _ in range(args.warmup):
        cpu_run(cfg, sample)
    t = []
    for _ in range(args.steps):
        secs, cores, kind = cpu_run(cfg, sample)
        t.append(secs)
    total = float(np.sum(t))
    value = units * args.steps / total
    line = dict(metric=metric_name(key), value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="u16 samples, f64 (+x87 long double) statistics", data="synthetic", impl="reference",
                config=dict(workload=cfg["name"], frames=cfg["n"], width=cfg["w"], height=cfg["h"], channels=cfg["c"],
                            sample=desc, same_config=rows == cfg["h"],
                            product_library_loaded="libsiril_stack_cuda" in open("/proc/self/maps").read()),
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
    ap.add_argument("--config", default="cfg3", choices=sorted(wl.CONFIGS))
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the end-to-end leg (default min(steps, 5))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: the library's own kernels over NVLink peer memory (default) or torch.distributed (NCCL) collectives")
    ap.add_argument("--small", action="store_true", help="1/16 of the frame area (debug only; not a bench value)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    key = args.config
    cfg = dict(wl.CONFIGS[key])
    if args.small:
        cfg.update(h=max(64, cfg["h"] // 4), w=max(64, (cfg["w"] // 4) & ~15), name=cfg["name"] + " [SMALL DEBUG SHAPE]")

    if args.impl == "reference":
        return reference_arm(args, key, cfg)

    import torch
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    pkg = ge.load_package()

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

    n, c, h, w, op = cfg["n"], cfg["c"], cfg["h"], cfg["w"], cfg["op"]
    norm, rej, sig, ro = cfg["normalize"], cfg["rejection"], cfg["sig"], cfg["row_order"]
    row0, rows = band_rows(h, world, rank)                  # this rank's RESULT rows
    src0 = h - (row0 + rows) if ro == 1 else row0           # ... come from these storage rows (SER: flipped)
    samples_total = n * c * h * w

    # ---- this rank's row band of every frame, generated in place (counter-based: any band alone) ----
    # (uint16 payloads live in int16 tensors: torch's uint16 kernels are incomplete; the library only sees pointers)
    frames = torch.empty((n, c, rows, w), dtype=torch.int16, device=dev)
    pkg.dev.synth_frames(frames.data_ptr(), n, c, rows, w, 0, src0, h, cfg["kind"], SEED, sp)
    out = torch.empty((c, rows, w), dtype=torch.int16, device=dev)
    band = pkg.dev.band(frames.data_ptr(), n, c, h, w, out_row0=row0, out_rows=rows, buf_row0=src0, buf_rows=rows,
                        row_order=ro)
    counters = torch.zeros(pkg.N_COUNTERS, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    coeff = torch.empty((3, n), dtype=torch.float64, device=dev)
    if norm:
        hist = torch.zeros((n, pkg.HIST_STRIDE), dtype=torch.int32, device=dev)
        scratch = torch.empty((n, 65536), dtype=torch.int32, device=dev)
        stats = torch.empty((4, n), dtype=torch.float64, device=dev)
        hband = pkg.dev.band(frames.data_ptr(), n, c, h, w, out_row0=src0, out_rows=rows, buf_row0=src0, buf_rows=rows)
    # ---- N > 1: the exchanges.  Default: the library's peer group -- every rank maps every other rank's block
    # (cudaIpc handles, all-gathered once here) and K4 / K5 / K6 exchange through it inside the kernels.
    group = None
    if world > 1 and args.exchange == "peer":
        group = pkg.PeerGroup(world, rank, n)
        mine = torch.frombuffer(bytearray(group.export()), dtype=torch.uint8).to(dev)
        allh = torch.empty((world, pkg.PEER_HANDLE_BYTES), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh.view(-1), mine)
        allh = allh.cpu().numpy()
        for r in range(world):
            group.import_(r, allh[r].tobytes())
        group.connect()
        counters_local = torch.zeros(pkg.N_COUNTERS, dtype=torch.int64, device=dev)
        dist.barrier()
    shard_stats = group is None and bool(norm) and world > 1 and n % world == 0   # reduce-scatter the histograms, IKSS on n / world frames per rank
    n_loc = n // world if shard_stats else n
    if shard_stats:
        hist_loc = torch.empty((n_loc, pkg.HIST_STRIDE), dtype=torch.int32, device=dev)
        stats_loc = torch.empty((4, n_loc), dtype=torch.float64, device=dev)
        stats_all = torch.empty((world, 4, n_loc), dtype=torch.float64, device=dev)
    if op == "sum":
        acc = torch.empty((c, rows, w), dtype=torch.int32, device=dev)
        maxim = torch.zeros(1, dtype=torch.int64, device=dev)
    # inputs smaller than twice the L2 are evicted between timed steps by writing a buffer larger than the L2
    flush = torch.empty(256 << 20, dtype=torch.int8, device=dev) if frames.numel() * 2 < 2 * L2_BYTES else None

    def step(ev=None):
        if ev: ev[0].record(stream)
        if op == "sum":
            maxim.zero_()
            pkg.dev.sum_accumulate(band, acc.data_ptr(), maxim.data_ptr(), sp)
            if ev: ev[1].record(stream)
            if group is not None:
                group.max_combine(maxim.data_ptr(), sp)            # stacking.c:311-313: one global max
            elif world > 1:
                dist.all_reduce(maxim, op=dist.ReduceOp.MAX)
            if ev: ev[2].record(stream)
            pkg.dev.sum_finish(acc.data_ptr(), c * rows * w, maxim.data_ptr(), out.data_ptr(), sp)
            if ev: ev[3].record(stream)
            return
        if op == "median":
            if ev: ev[1].record(stream); ev[2].record(stream)
            pkg.dev.stack_median(band, out.data_ptr(), stream=sp)
            if ev: ev[3].record(stream)
            return
        counters.zero_()
        if norm and group is not None:
            # K4 adds this band's counts straight into the owners' histograms (frame f: rank f % world), K5 runs on
            # the owned frames once all bands arrived and stores the statistics into every rank's table, K6 waits
            # for all owners: no collective call, no host in between (csrc/ssc_peer.cu)
            group.histogram_scatter(hband, True, sp)
            if ev: ev[1].record(stream)
            group.ikss_gather(n, scratch.data_ptr(), sp)
            group.norm_coeffs_gathered(n, 0, norm, coeff.data_ptr(), status.data_ptr(), None, sp)
        elif norm:
            hist.zero_()
            pkg.dev.histogram(hband, hist.data_ptr(), sp)
            if shard_stats:
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
            pkg.dev.norm_coeffs(stats.data_ptr(), n, 0, norm, coeff.data_ptr(), status.data_ptr(), sp)
        elif ev:
            ev[1].record(stream)
        if ev: ev[2].record(stream)
        if group is not None:
            counters_local.zero_()
            pkg.dev.stack_mean(band, out.data_ptr(), counters_local.data_ptr(), rej, sig, norm, coeff.data_ptr() if norm else None,
                               stream=sp)
            if ev: ev[3].record(stream)
            group.counters_combine(counters_local.data_ptr(), counters.data_ptr(), sp)
            return
        pkg.dev.stack_mean(band, out.data_ptr(), counters.data_ptr(), rej, sig, norm, coeff.data_ptr() if norm else None,
                           stream=sp)
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
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 1)                              # > L2: nothing of the previous step stays cached
        step(evs[k])
        ends[k].record(stream)
    sync_all()
    ms_total = float(sum(evs[k][0].elapsed_time(ends[k]) for k in range(args.steps)))   # flushes excluded
    phase = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in evs])
    tt = torch.tensor([ms_total] + phase.mean(0).tolist(), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)           # max over ranks
    ms_total, ms_p0, ms_p1, ms_p2 = tt.tolist()
    ms_step = ms_total / args.steps
    value = samples_total / 1e9 / (ms_step / 1e3)
    assert int(status.item()) == 0
    irej = counters[:6].cpu().numpy().astype(np.uint64).reshape(3, 2)
    guard, cap = int(counters[6].item()), int(counters[7].item())

    # ---- end to end: pinned host frames -> result on the host, copies inside the timer ----
    # One operator call on the whole image, as a Siril host makes it.  N > 1: ssc_set_devices(0..N-1) and the SAME
    # single call from rank 0 -- the library splits the image over the GPUs itself (one host thread and three
    # streams per device, NVLink exchanges between the devices); the other ranks only hold the barrier.
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 5)
        h2d_bytes = n * c * h * w * 2
        d2h_bytes = c * h * w * 2
        e2e_ms, r = 0.0, None
        e2e_api = "ssc_stack_%s (operator layer, whole image%s)" % (
            dict(sum="summing", median="median", mean="mean_with_rejection")[op],
            "" if world == 1 else ", ssc_set_devices(0..%d), one call from rank 0" % (world - 1))
        if rank == 0:
            host = torch.empty((n, c, h, w), dtype=torch.int16, pin_memory=True)
            if world == 1:
                host.copy_(frames)
            else:                                           # the other bands, regenerated here (counter-based generator)
                for rr in range(world):
                    b0, br = band_rows(h, world, rr)
                    s0 = h - (b0 + br) if ro == 1 else b0
                    tmp = torch.empty((n, c, br, w), dtype=torch.int16, device=dev)
                    pkg.dev.synth_frames(tmp.data_ptr(), n, c, br, w, 0, s0, h, cfg["kind"], SEED, sp)
                    host[:, :, s0:s0 + br, :].copy_(tmp)
                    del tmp
            torch.cuda.synchronize()
            seq = pkg.Sequence([host[i].numpy().view(np.uint16) for i in range(n)], row_order=ro)
            pkg.set_devices(list(range(world)))

            def e2e_step():
                if op == "sum":
                    return pkg.stack_summing(seq)
                if op == "median":
                    return pkg.stack_median(seq)
                return pkg.stack_mean_with_rejection(seq, rej, sig, norm, want_coeff=False)
            for _ in range(2):
                r = e2e_step()
            t_wall0 = time.perf_counter()
            for _ in range(e2e_steps):
                r = e2e_step()
            e2e_ms = (time.perf_counter() - t_wall0) * 1e3 / max(e2e_steps, 1)
            pkg.set_devices([0])
        # the other ranks wait on the HOST (the store): a NCCL barrier would spin a kernel on their GPUs and time-slice
        # with the kernels rank 0 runs there
        if world > 1:
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                store.set("ssc_e2e_done", "1")
            else:
                store.wait(["ssc_e2e_done"])
        sync_all()
        e2e_parity, op_timing = None, None
        if rank == 0 and r is not None:
            mine = r.data[:, row0:row0 + rows, :]           # this rank's band of the operator's image
            e2e_parity = bool(np.array_equal(mine, out.cpu().numpy().view(np.uint16)))
            if op == "mean":
                e2e_parity = e2e_parity and bool(np.array_equal(r.irej, irej))
            op_timing = dict(h2d_ms=r.t_h2d_ms, tail_after_upload_ms=r.t_device_ms, total_ms=r.t_total_ms)
            e2e = dict(value=samples_total / 1e9 / (e2e_ms / 1e3), unit=UNIT, h2d_bytes_per_step=h2d_bytes,
                       d2h_bytes_per_step=d2h_bytes, ms_per_step=e2e_ms, steps=e2e_steps, api=e2e_api,
                       operator_timing_ms=op_timing, matches_device_path=e2e_parity)
            del host, seq, r
            pkg.release_workspace()

    clocks = sampler.stop() if sampler else None

    # ---- parity spot check of this very run against the oracle (not timed, small sample) ----
    parity = None
    if rank == 0:
        try:
            from oracle import port
            ys = np.linspace(0, rows - 1, 6).astype(int)               # result rows of this band
            src = (rows - 1 - ys) if ro == 1 else ys                   # their storage rows in the band buffer
            sub = frames.index_select(2, torch.as_tensor(src.copy(), device=dev)).cpu().numpy().view(np.uint16)
            got = out.index_select(1, torch.as_tensor(ys, device=dev)).cpu().numpy().view(np.uint16)
            if op == "sum":
                # per-pixel sums of the sampled rows rescaled with the run's global max (stacking.c:328-341)
                m = int(maxim.item())
                s = sub.astype(np.int64).sum(0)
                if row0 == 0 and ro == 0:
                    s[:, 0, 0] = 0                                      # the `ii > 0` quirk (stacking.c:307)
                exp = np.floor(s * (65535.0 / m) + 0.5).clip(0, 65535) if m > 65535 else s
                parity = dict(rows_checked=len(ys), bit_exact=bool(np.array_equal(got, exp.astype(np.uint16))))
            elif op == "median":
                parity = dict(rows_checked=len(ys), bit_exact=bool(np.array_equal(got, port.stack_median(sub)["out"])))
            else:
                cf = coeff.cpu().numpy() if norm else None
                chk = port.stack_mean(sub, rej, sig, norm, coeff=cf, sd_mode=port.SD_EXACT)
                gsl = port.stack_mean(sub, rej, sig, norm, coeff=cf, sd_mode=port.SD_GSL)
                d = np.abs(got.astype(np.int32) - gsl["out"].astype(np.int32))
                parity = dict(rows_checked=len(ys), bit_exact_vs_oracle_same_sigma_form=bool(np.array_equal(got, chk["out"])),
                              max_abs_diff_vs_reference_form=int(d.max()), frac_identical_vs_reference_form=float((d == 0).mean()))
        except Exception as e:                                           # the checker is optional at bench time
            parity = dict(error=repr(e))

    # ---- CPU baseline: the reference's own code on this box's cores, bounded sample (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            crow = cpu_sample_rows(cfg, 12.0, 1)
            c0 = (h - crow) // 2
            sample = wl.host_frames(cfg, c0, crow)
            secs, cores, kind = cpu_run(cfg, sample)
            units, desc = sample_desc(cfg, c0, crow)
            cpu = dict(value=units / secs, unit=UNIT, cores=cores, kind=kind, seconds=secs, sample=desc + ", one pass")
        except Exception as e:
            cpu = dict(error=repr(e))

    if rank == 0:
        peak, peak_src = measured_peaks()
        rows_max = band_rows(h, world, 0)[1]
        bytes_stack, bytes_stats = wl.algorithmic_bytes(cfg, rows_max)  # per launch, on the largest band
        if op == "sum":
            kernel_ms, kernel_bytes = ms_p0, 2 * n * c * rows_max * w     # K1 reads every sample once; K1b writes the result
            breakdown = dict(accumulate=ms_p0, max_exchange=ms_p1, rescale=ms_p2)
            launches = 2
        elif op == "median":
            kernel_ms, kernel_bytes = ms_p2, bytes_stack
            breakdown = dict(stack=ms_p2)
            launches = 1
        else:
            kernel_ms, kernel_bytes = ms_p2, bytes_stack
            breakdown = dict(histogram_allreduce=ms_p0, ikss_coeffs=ms_p1, stack=ms_p2)
            if norm:
                breakdown["histogram_gbs"] = bytes_stats / (ms_p0 / 1e3) / 1e9
            launches = (4 if c > 1 else 3) * bool(norm) + 2             # K4 (+ other-layer check), K5, K6; K3 (+ hand-backs / tail)
        achieved = kernel_bytes / (kernel_ms / 1e3) / 1e9
        full_stack, full_stats = wl.algorithmic_bytes(cfg)
        traffic, traffic_src = DRAM_BYTES_N1.get(key, (None, None))
        in_bytes = 2 * samples_total
        line = dict(
            metric=metric_name(key), value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_step, higher_is_better=True, scaling="strong", vs_baseline=None,
            dtype="u16 samples, f64 clipping/normalization arithmetic" if op == "mean" else "u16", data="synthetic",
            config=dict(workload=cfg["name"], frames=n, width=w, height=h, channels=c,
                        parallelism="row bands x%d" % world,
                        exchange=(None if world == 1 else ("libsiril_stack_cuda peer group (kernels over NVLink peer memory)"
                                                           if group is not None else "torch.distributed NCCL collectives")),
                        l2=("inputs (%.1f GB) exceed L2 (126 MB): no flush needed" % (in_bytes / 1e9) if flush is None else
                            "inputs (%.0f MB) fit L2: a 256 MB buffer is written between timed steps (outside the per-step events)" % (in_bytes / world / 1e6)),
                        seed=hex(SEED), rejection_counters=irej.tolist(), guard_fired=guard, cap_hit=cap),
            roofline=dict(bound="hbm", kernel=DOMINANT[key], achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=(traffic // world) if traffic else None,
                          traffic_source=("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture: %s; / ranks" % traffic_src) if traffic else None,
                          peak_source=peak_src, algorithmic_bytes_per_launch=kernel_bytes, ms_per_launch=kernel_ms),
            step_breakdown_ms=breakdown,
            step_hbm_frac=(full_stack + full_stats) / world / (ms_step / 1e3) / 1e9 / peak,
            e2e=e2e, gpu_launches=launches * args.steps, clocks=clocks, parity=parity, cpu_baseline=cpu)
        _emit(line)
    if group is not None:
        assert group.status(sp) == 0, "a peer exchange timed out"
        dist.barrier()
        group.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
