"""Operator-layer paths that only large inputs reach on their own: the HBM-budget row split
(several pieces per device, SSC_HBM_BUDGET_BYTES forces it; the reference's counterpart is the
max_number_of_rows block split, stacking.c:570-646,1391-1476) and several devices in one operator
call (ssc_set_devices)."""
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


class budget:
    """Per-call HBM budget for frames + result, in bytes (read by the operator layer at every call)."""

    def __init__(self, nbytes):
        self.nbytes = nbytes

    def __enter__(self):
        os.environ["SSC_HBM_BUDGET_BYTES"] = str(int(self.nbytes))

    def __exit__(self, *a):
        os.environ.pop("SSC_HBM_BUDGET_BYTES", None)


def per_row_bytes(n, c, w):
    return n * c * w * 2 + c * w * 10          # what ssc_api.cu books per result row


SX = [0, 3, -2, 1, 0, -4, 2, 0, 1, -1, 0, 2]
SY = [0, -3, 2, 4, -1, 0, 1, -2, 3, 0, -4, 1]


@pytest.mark.parametrize("pieces", [2, 3, 7])
@pytest.mark.parametrize("row_order", [0, 1])
def test_row_split_sum_median_extremes(gpu, port, pieces, row_order):
    n, c, h, w = 12, 3, 61, 40
    fr = synth.make_random(n, c, h, w, seed=3, hi=9000)
    rows_fit = (h + pieces - 1) // pieces
    plain = gpu.Sequence(fr, row_order=row_order)
    shifted = gpu.Sequence(fr, row_order=row_order, shiftx=SX, shifty=SY)
    with budget(rows_fit * per_row_bytes(n, c, w)):
        for seq, kw in ((plain, {}), (shifted, dict(shiftx=SX, shifty=SY))):
            res = gpu.stack_summing(seq)
            p = port.stack_sum(fr, row_order=row_order, **kw)
            assert np.array_equal(res.data, p["out"])                   # the global max spans all pieces
            assert res.hi == min(p["maxim"], 65535)
            for take_max in (True, False):
                res = (gpu.stack_addmax if take_max else gpu.stack_addmin)(seq)
                p = port.stack_extreme(fr, take_max, row_order=row_order, **kw)
                assert np.array_equal(res.data, p["out"]) and res.hi == p["hi"]
        res = gpu.stack_median(plain)
        assert np.array_equal(res.data, port.stack_median(fr, row_order=row_order)["out"])


@pytest.mark.parametrize("pieces", [3, 5])
@pytest.mark.parametrize("rej,sig", [(2, (4.0, 3.0)), (4, (4.0, 3.0)), (5, (5.0, 2.5))])
def test_row_split_mean_with_normalization(gpu, port, pieces, rej, sig):
    """Normalized mean over several pieces: pass 1 histograms every piece, pass 2 uploads them again
    and stacks; result, counters and coefficients equal the unsplit run and the oracle."""
    n, c, h, w = 12, 3, 64, 40
    fr = synth.make_frames(n, c, h, w, seed=21, outlier_p=5e-3, zero_p=1e-3, stars=4)
    rows_fit = (h + pieces - 1) // pieces
    for kw in ({}, dict(shiftx=SX, shifty=SY)):
        for ro in (0, 1):
            seq = gpu.Sequence(fr, row_order=ro, **kw)
            whole = gpu.stack_mean_with_rejection(seq, rej, sig, 3)
            # shifts widen every piece's source rows by their reach: budget for them too
            reach = (max(SY) - min(SY)) if kw else 0
            with budget((rows_fit + 1) * per_row_bytes(n, c, w) + reach * n * c * w * 2):
                part = gpu.stack_mean_with_rejection(seq, rej, sig, 3)
            assert np.array_equal(part.coeff, whole.coeff)
            assert np.array_equal(part.data, whole.data), (ro, bool(kw))
            assert np.array_equal(part.irej, whole.irej)
            same = port.stack_mean(fr, rej, sig, 3, coeff=whole.coeff, row_order=ro, sd_mode=port.SD_EXACT, **kw)
            assert np.array_equal(part.data, same["out"])
            assert np.array_equal(part.irej, same["irej"])


def test_row_split_really_splits(gpu):
    """The budget must change the plan (phase timings show several uploads), and a budget below one
    row is an error, not a crash."""
    n, c, h, w = 8, 1, 48, 64
    fr = synth.make_random(n, c, h, w, seed=9)
    seq = gpu.Sequence(fr)
    logs = []
    gpu.set_callbacks(progress=lambda t, f: logs.append(f))
    try:
        with budget(10 * per_row_bytes(n, c, w)):
            gpu.stack_median(seq)
    finally:
        gpu.set_callbacks()
    assert len([f for f in logs if 0 < f <= 1.0]) >= 5               # one progress step per piece
    with budget(per_row_bytes(n, c, w) // 2):
        with pytest.raises(gpu.StackError) as e:
            gpu.stack_median(seq)
    assert e.value.code == -1


def test_set_devices_rejects_duplicates_and_missing(gpu):
    with pytest.raises(gpu.StackError) as e:
        gpu.set_devices([0, 0])
    assert e.value.code == gpu.ERR_ARG
    with pytest.raises(gpu.StackError) as e:
        gpu.set_devices([gpu.device_count()])
    assert e.value.code == gpu.ERR_CUDA
    gpu.set_devices([0])


def test_default_normalization_label_of_the_mean(gpu, port):
    """An out-of-range args->normalize: stack_mean_with_rejection's own `default:` label stacks the raw
    samples (stacking.c:1636-1638), stack_median's normalizes additively with scale (:753-757)."""
    fr = synth.make_frames(10, 1, 24, 32, seed=4)
    seq = gpu.Sequence(fr)
    res = gpu.stack_mean_with_rejection(seq, 2, (4.0, 3.0), normalize=9, want_coeff=False)
    same = port.stack_mean(fr, 2, (4.0, 3.0), 0, sd_mode=port.SD_EXACT)
    assert np.array_equal(res.data, same["out"]) and np.array_equal(res.irej, same["irej"])
    coeff = gpu.compute_normalization(seq, 3)
    res = gpu.stack_median(seq, normalize=9)
    assert np.array_equal(res.data, port.stack_median(fr, normalize=3, coeff=coeff)["out"])


# ------------------------------------------------------------------------------ several devices
def multi():
    import __graft_entry__ as g
    return g.load_package().device_count() >= 2


@pytest.mark.skipif("not multi()", reason="needs two CUDA devices (run with gpurun --gpus 2)")
@pytest.mark.parametrize("ndev", [2, 3, 8])
def test_several_devices_in_one_operator_call(gpu, port, ndev):
    """ssc_set_devices: the image is split in row bands over the devices inside one stack_method call;
    histograms are combined over NVLink peer access (exact integer sums) before the IKSS."""
    ndev = min(ndev, gpu.device_count())
    n, c, h, w = 16, 3, 53, 48
    fr = synth.make_frames(n, c, h, w, seed=8, outlier_p=5e-3, zero_p=1e-3, stars=4)
    rnd = synth.make_random(n, c, h, w, seed=5, hi=9000)
    sx, sy = SX + [0, 1, -1, 2], SY + [2, 0, -1, 1]
    try:
        for ro in (0, 1):
            for kw in ({}, dict(shiftx=sx, shifty=sy)):
                gpu.set_devices([0])
                s1, r1 = gpu.Sequence(fr, row_order=ro, **kw), gpu.Sequence(rnd, row_order=ro, **kw)
                one = [gpu.stack_mean_with_rejection(s1, 2, (4.0, 3.0), 2), gpu.stack_mean_with_rejection(s1, 4, (4.0, 3.0), 3),
                       gpu.stack_median(s1, normalize=1), gpu.stack_summing(r1), gpu.stack_addmax(r1), gpu.stack_addmin(r1)]
                gpu.set_devices(list(range(ndev)))
                many = [gpu.stack_mean_with_rejection(s1, 2, (4.0, 3.0), 2), gpu.stack_mean_with_rejection(s1, 4, (4.0, 3.0), 3),
                        gpu.stack_median(s1, normalize=1), gpu.stack_summing(r1), gpu.stack_addmax(r1), gpu.stack_addmin(r1)]
                for a, b in zip(one, many):
                    assert np.array_equal(a.data, b.data), (ro, bool(kw))
                    assert np.array_equal(a.irej, b.irej) and a.hi == b.hi
                    if a.coeff is not None:
                        assert np.array_equal(a.coeff, b.coeff)         # statistics exact under sharding
                # and with a row split inside every device's band
                with budget(5 * per_row_bytes(n, c, w) + 8 * n * c * w * 2):
                    b = gpu.stack_mean_with_rejection(s1, 2, (4.0, 3.0), 2)
                assert np.array_equal(one[0].data, b.data) and np.array_equal(one[0].irej, b.irej)
        same = port.stack_mean(fr, 2, (4.0, 3.0), 2, coeff=many[0].coeff, row_order=1, shiftx=sx, shifty=sy,
                               sd_mode=port.SD_EXACT)
        assert np.array_equal(many[0].data, same["out"])
    finally:
        gpu.set_devices([0])


@pytest.mark.skipif("not multi()", reason="needs two CUDA devices (run with gpurun --gpus 2)")
@pytest.mark.parametrize("n", [5, 16])
def test_peer_group_exchange_in_one_process(gpu, port, n):
    """ssc_peer_*: the NVLink exchanges of the device layer driven from one process (ssc_peer_attach_local).
    Every device holds a row band; K4 adds its counts into the owners' histograms, K5 runs on the owned
    frames, the statistics are stored into every device's table, K6 follows: coefficients, image and
    counters equal the single-device pipeline bit for bit -- also on a second call (epochs), with a frame
    count that does not divide by the device count, and for stack_summing's maximum."""
    import torch
    from test_gpu_device_layer import dev_u16, host_u16, run_pipeline
    ndev = min(gpu.device_count(), 4)
    c, h, w = 3, 41, 48
    fr = synth.make_frames(n, c, h, w, seed=40 + n, outlier_p=5e-3, zero_p=1e-3, stars=3)
    torch.cuda.set_device(0)
    ref_out, ref_cnt, ref_coeff = run_pipeline(gpu, dev_u16(fr), n, c, h, w, [(0, h)], 2, (4.0, 3.0), 3)
    groups = []
    for g in range(ndev):
        torch.cuda.set_device(g)
        groups.append(gpu.PeerGroup(ndev, g, n))
    for g in range(ndev):
        torch.cuda.set_device(g)
        for o in range(ndev):
            groups[g].attach_local(o, groups[o])
        groups[g].connect()
    rows = [(g * h // ndev, (g + 1) * h // ndev - g * h // ndev) for g in range(ndev)]
    bufs = []
    for g in range(ndev):
        torch.cuda.set_device(g)
        dv = torch.device("cuda", g)
        bufs.append(dict(frames=torch.from_numpy(fr.view(np.int16)).to(dv), out=torch.zeros((c, h, w), dtype=torch.int16, device=dv),
                         scratch=torch.empty(((n + ndev - 1) // ndev, 65536), dtype=torch.int32, device=dv),
                         coeff=torch.empty((3, n), dtype=torch.float64, device=dv), status=torch.zeros(1, dtype=torch.int32, device=dv),
                         cnt=torch.zeros(8, dtype=torch.int64, device=dv), tot=torch.zeros(8, dtype=torch.int64, device=dv),
                         mx=torch.zeros(1, dtype=torch.int64, device=dv)))
    try:
        for call in range(2):
            for stage in range(4):                      # every stage on all devices before the next: nothing may block the host
                for g in range(ndev):
                    torch.cuda.set_device(g)
                    b, sp = bufs[g], torch.cuda.current_stream().cuda_stream
                    r0, rn = rows[g]
                    band = gpu.dev.band(b["frames"].data_ptr(), n, c, h, w, out_row0=r0, out_rows=rn)
                    if stage == 0:
                        b["cnt"].zero_()
                        groups[g].histogram_scatter(band, True, sp)
                    elif stage == 1:
                        groups[g].ikss_gather(n, b["scratch"].data_ptr(), sp)
                    elif stage == 2:
                        groups[g].norm_coeffs_gathered(n, 0, 3, b["coeff"].data_ptr(), b["status"].data_ptr(), None, sp)
                        gpu.dev.stack_mean(band, b["out"].data_ptr(), b["cnt"].data_ptr(), 2, (4.0, 3.0), 3, b["coeff"].data_ptr(),
                                           out_layer_stride=h * w, out_buf_row0=0, stream=sp)
                    else:
                        groups[g].counters_combine(b["cnt"].data_ptr(), b["tot"].data_ptr(), sp)
                        b["mx"].fill_(1000 * (g + 1) + call)
                        groups[g].max_combine(b["mx"].data_ptr(), sp)
            out = np.zeros((c, h, w), np.uint16)
            for g in range(ndev):
                torch.cuda.set_device(g)
                torch.cuda.synchronize()
                assert groups[g].status() == 0
                r0, rn = rows[g]
                out[:, r0:r0 + rn] = host_u16(bufs[g]["out"])[:, r0:r0 + rn]
                assert np.array_equal(bufs[g]["coeff"].cpu().numpy(), ref_coeff)         # exact under sharding
                assert np.array_equal(bufs[g]["tot"].cpu().numpy(), ref_cnt)             # the same totals everywhere
                assert int(bufs[g]["mx"].item()) == 1000 * ndev + call
            assert np.array_equal(out, ref_out)
    finally:
        for g in range(ndev):
            torch.cuda.set_device(g)
            groups[g].close()
        torch.cuda.set_device(0)


@pytest.mark.parametrize("row_order", [0, 1])
def test_chunked_pipeline_equals_one_chunk(gpu, port, row_order, monkeypatch):
    """The operators upload, stack and download a resident band in row chunks per layer (layer 0 first when
    statistics are needed: K4 runs behind the copy front).  SSC_CHUNK_BYTES forces many small chunks: every
    operator must return what it returns in one chunk, and what the oracle says -- also for frames that
    are not equally spaced in host memory (per-frame copies instead of one pitched copy) and for an 8-bit
    sequence whose other layers decide the histogram range (utils.c:454-459)."""
    n, c, h, w = 10, 3, 37, 40
    fr = synth.make_frames(n, c, h, w, seed=77, outlier_p=5e-3, zero_p=1e-3, stars=3)
    rnd = synth.make_random(n, c, h, w, seed=6, hi=9000)
    scattered = [fr[i].copy() for i in (3, 0, 7, 1, 9, 2, 8, 4, 6, 5)]          # separate allocations, shuffled
    small = synth.make_random(4, 3, 16, 32, seed=1, lo=0, hi=200)
    small[1, 2, 3, 3] = 300                                                    # another layer lifts one frame's max
    def run_all():
        s, r = gpu.Sequence(fr, row_order=row_order), gpu.Sequence(rnd, row_order=row_order)
        return [gpu.stack_mean_with_rejection(s, 2, (4.0, 3.0), 2), gpu.stack_mean_with_rejection(s, 4, (4.0, 3.0), 3),
                gpu.stack_mean_with_rejection(s, 5, (5.0, 2.5), 0), gpu.stack_median(s, normalize=1, want_coeff=True),
                gpu.stack_summing(r), gpu.stack_addmax(r), gpu.stack_addmin(r),
                gpu.stack_mean_with_rejection(gpu.Sequence(scattered, row_order=row_order), 2, (4.0, 3.0), 4),
                gpu.stack_mean_with_rejection(gpu.Sequence(small, row_order=row_order), 0, (0, 0), 3)]
    one = run_all()
    monkeypatch.setenv("SSC_CHUNK_BYTES", str(5 * n * w * 2))                  # five rows per chunk
    many = run_all()
    monkeypatch.delenv("SSC_CHUNK_BYTES")
    for a, b in zip(one, many):
        assert np.array_equal(a.data, b.data) and np.array_equal(a.irej, b.irej) and a.hi == b.hi
        assert a.guard_fired == b.guard_fired
        if a.coeff is not None:
            assert np.array_equal(a.coeff, b.coeff)
    same = port.stack_mean(fr, 2, (4.0, 3.0), 2, coeff=many[0].coeff, row_order=row_order, sd_mode=port.SD_EXACT)
    assert np.array_equal(many[0].data, same["out"]) and np.array_equal(many[0].irej, same["irej"])
    assert np.array_equal(many[4].data, port.stack_sum(rnd, row_order=row_order)["out"])
    st = [port.frame_stats_sorted(f) for f in small]
    assert st[1][2] == 65535.0 and st[0][2] == 255.0
    assert np.allclose(many[8].coeff, port.norm_coeffs(3, [x[0] for x in st], [x[1] for x in st]), rtol=1e-11)
