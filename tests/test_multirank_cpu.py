"""The N>1 path on CPU: world_size-2 `gloo` processes shard the image in row bands exactly as
bench.py / the library do (remainder rows to the first ranks, stacking.c:620-625), all-reduce the
exact per-frame histograms (integer SUM) and must end up with the coefficients and the stacked
rows of the unsharded run.  The per-band compute is the oracle here (no GPU in this test): what is
under test is the sharding arithmetic and that the statistics are exact under any partition."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_band_rows_partition():
    bench = _bench_module()
    for h in (1, 7, 152, 1216, 4000, 4001):
        for world in (1, 2, 3, 4, 8):
            spans = [bench.band_rows(h, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (a0, an), (b0, bn) in zip(spans, spans[1:]):
                assert a0 + an == b0                     # contiguous, no overlap
            assert spans[-1][0] + spans[-1][1] == h
            sizes = [s[1] for s in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _worker(rank, world, port_no, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from oracle import port
    bench = _bench_module()
    n, c, h, w = 10, 3, 37, 48
    frames = synth.make_frames(n, c, h, w, seed=5, outlier_p=5e-3, zero_p=2e-3, stars=3)   # same on every rank
    row0, rows = bench.band_rows(h, world, rank)
    band = np.ascontiguousarray(frames[:, :, row0:row0 + rows, :])
    # K4 on the band, all-reduce(SUM), K5/K6 on every rank
    hist = np.stack([port.histogram(f)[0] for f in band]).astype(np.int64)
    t = torch.from_numpy(hist)
    dist.all_reduce(t)
    hist = t.numpy().astype(np.uint32)
    hist[:, 0] = 0
    stats = [port.ikss_hist(hk) for hk in hist]
    coeff = port.norm_coeffs(port.MULTIPLICATIVE_SCALING, [s[0] for s in stats], [s[1] for s in stats])
    out = port.stack_mean(band, port.SIGMA, (4.0, 3.0), port.MULTIPLICATIVE_SCALING, coeff=coeff, sd_mode=port.SD_EXACT)
    cnt = torch.from_numpy(out["irej"].astype(np.int64).copy())
    dist.all_reduce(cnt)
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), out=out["out"], coeff=coeff, irej=cnt.numpy(), row0=row0, rows=rows)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(tmp_path, world):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)

    import synth
    from oracle import port
    n, c, h, w = 10, 3, 37, 48
    frames = synth.make_frames(n, c, h, w, seed=5, outlier_p=5e-3, zero_p=2e-3, stars=3)
    stats = [port.frame_stats_hist(f) for f in frames]
    coeff = port.norm_coeffs(port.MULTIPLICATIVE_SCALING, [s[0] for s in stats], [s[1] for s in stats])
    whole = port.stack_mean(frames, port.SIGMA, (4.0, 3.0), port.MULTIPLICATIVE_SCALING, coeff=coeff, sd_mode=port.SD_EXACT)
    got = np.zeros_like(whole["out"])
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(z["coeff"], coeff)                        # exact under sharding, on every rank
        assert np.array_equal(z["irej"].reshape(3, 2), whole["irej"].astype(np.int64))
        got[:, int(z["row0"]):int(z["row0"]) + int(z["rows"]), :] = z["out"]
    assert np.array_equal(got, whole["out"])


def _worker_owners(rank, world, port_no, tmpdir):
    """The peer-group protocol of csrc/ssc_peer.cu restated on gloo: every rank counts its band, frame f is
    analysed by rank f % world from the sum of all ranks' counts (slot s of a rank is frame s * world + rank),
    the four statistics per frame are gathered on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from oracle import port
    bench = _bench_module()
    n, c, h, w = 7, 1, 29, 40                                           # 7 frames: not a multiple of 2 or 3
    frames = synth.make_frames(n, c, h, w, seed=9, outlier_p=5e-3, zero_p=2e-3, stars=3)
    row0, rows = bench.band_rows(h, world, rank)
    band = np.ascontiguousarray(frames[:, :, row0:row0 + rows, :])
    mine = torch.from_numpy(np.stack([port.histogram(f)[0] for f in band]).astype(np.int64))
    everyone = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)                                     # (the kernels read the peers' blocks instead)
    owned = (n - rank + world - 1) // world
    own_stats = torch.zeros((4, max(owned, 1)), dtype=torch.float64)
    for s in range(owned):
        f = s * world + rank
        hk = sum(e[f] for e in everyone).numpy().astype(np.uint32)
        hk[0] = 0
        loc, sc, _ = port.ikss_hist(hk)
        own_stats[:, s] = torch.tensor([loc, sc, 65535.0, float(hk.sum())], dtype=torch.float64)
    pad = torch.zeros((4, (n + world - 1) // world), dtype=torch.float64)
    pad[:, :own_stats.shape[1]] = own_stats[:, :pad.shape[1]]
    table = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(table, pad)
    stats = np.zeros((4, n))
    for r in range(world):
        for s in range(pad.shape[1]):
            if s * world + r < n:
                stats[:, s * world + r] = table[r][:, s].numpy()
    coeff = port.norm_coeffs(port.ADDITIVE_SCALING, stats[0], stats[1])
    np.savez(os.path.join(tmpdir, "own%d.npz" % rank), coeff=coeff, stats=stats)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_owners_protocol(tmp_path, world):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_worker_owners, args=(world, port_no, str(tmp_path)), nprocs=world, join=True)
    import synth
    from oracle import port
    frames = synth.make_frames(7, 1, 29, 40, seed=9, outlier_p=5e-3, zero_p=2e-3, stars=3)
    st = [port.frame_stats_hist(f) for f in frames]
    coeff = port.norm_coeffs(port.ADDITIVE_SCALING, [x[0] for x in st], [x[1] for x in st])
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "own%d.npz" % r))
        assert np.array_equal(z["coeff"], coeff)                        # every rank, bit for bit
        assert np.array_equal(z["stats"][0], [x[0] for x in st]) and np.array_equal(z["stats"][1], [x[1] for x in st])
