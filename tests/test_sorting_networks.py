"""CPU checks of the compare-exchange schedules the CUDA kernels execute (no GPU needed):
the generated register networks (csrc/ssc_sortnet.h) and the shared-memory merge schedule of
csrc/ssc_stack_mid.cu, restated index for index."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "siril-0.9_b200", "csrc", "ssc_sortnet.h")


def networks():
    text = open(HEADER).read()
    nets = {}
    for m in re.finditer(r"#define SSC_SORTNET_(\d+) ((?:.*\\\n)*.*\n)", text):
        nets[int(m.group(1))] = [(int(a), int(b)) for a, b in re.findall(r"SSC_CE\((\d+),(\d+)\)", m.group(2))]
    return nets


def apply_network(ces, v):
    v = v.copy()
    for a, b in ces:                     # SSC_CE(a, b): min to v[a], max to v[b]; vectorised over the columns of v
        lo, hi = np.minimum(v[a], v[b]), np.maximum(v[a], v[b])
        v[a], v[b] = lo, hi
    return v


def pruned_networks():
    """SSC_SORTNET_<n>_SIGMA: [(op, a, b)] with op in CE / MIN / MAX."""
    text = open(HEADER).read()
    nets = {}
    for m in re.finditer(r"#define SSC_SORTNET_(\d+)_SIGMA ((?:.*\\\n)*.*\n)", text):
        nets[int(m.group(1))] = [(op, int(a), int(b)) for op, a, b in re.findall(r"SSC_(CE|MIN|MAX)\((\d+),(\d+)\)", m.group(2))]
    return nets


@pytest.mark.parametrize("n", [32, 64])
def test_pruned_sigma_networks_place_the_ranks_the_kernel_reads(n):
    """The register sigma clipping (ssc_stack_fast.cu) reads ranks 0..2, n/2-3..n/2+1 and n-6..n-1 only; the
    pruned network (dead comparator halves removed, tools/gen_sortnet.py) must leave exactly the sorted
    values there: all threshold (0-1) inputs in random order, heavy ties, and random u16 columns."""
    ops = pruned_networks()[n]
    full = networks()[n]
    n_ops = sum(2 if op == "CE" else 1 for op, _, _ in ops)
    assert n_ops == {32: 332, 64: 888}[n] and n_ops < 2 * len(full)
    ranks = [0, 1, 2] + list(range(n // 2 - 3, n // 2 + 2)) + list(range(n - 6, n))
    rng = np.random.default_rng(n)
    cols = [rng.permutation(np.r_[np.zeros(n - t, np.int64), np.ones(t, np.int64)]) for t in range(n + 1) for _ in range(40)]
    cols += [rng.integers(0, 4, n) for _ in range(500)] + [rng.integers(0, 65536, n) for _ in range(2000)]
    v = np.array(cols).T.copy()                                       # [n, columns]
    want = np.sort(v, axis=0)
    for op, a, b in ops:
        lo, hi = np.minimum(v[a], v[b]), np.maximum(v[a], v[b])
        if op in ("CE", "MIN"):
            v[a] = lo
        if op in ("CE", "MAX"):
            v[b] = hi
    assert np.array_equal(v[ranks], want[ranks])


def test_header_has_the_three_networks():
    nets = networks()
    assert sorted(nets) == [16, 32, 64]
    assert [len(nets[n]) for n in (16, 32, 64)] == [63, 191, 543]      # Batcher odd-even merge sort
    for n, ces in nets.items():
        assert all(0 <= a < b < n for a, b in ces)


def test_network_16_sorts_every_0_1_input():
    """0-1 principle: a comparator network that sorts all 2^n binary inputs sorts everything."""
    ces = networks()[16]
    cols = ((np.arange(1 << 16, dtype=np.uint32)[None, :] >> np.arange(16, dtype=np.uint32)[:, None]) & 1).astype(np.uint8)
    out = apply_network(ces, cols)
    assert (np.diff(out.astype(np.int8), axis=0) >= 0).all()


@pytest.mark.parametrize("n", [32, 64])
def test_networks_32_64_sort_random_and_structured_inputs(n):
    ces = networks()[n]
    rng = np.random.default_rng(n)
    cols = [rng.integers(0, 65536, size=(n, 20000)), rng.integers(0, 3, size=(n, 20000)),
            rng.integers(0, 2, size=(n, 200000)),
            np.stack([np.roll(np.arange(n), k) for k in range(n)], axis=1),
            np.stack([np.roll(np.arange(n)[::-1], k) for k in range(n)], axis=1)]
    # every 0-1 input with a single run of ones anywhere (rotations of sorted 0-1 vectors)
    runs = np.array([[1 if (i - s) % n < k else 0 for i in range(n)] for s in range(n) for k in range(n + 1)]).T
    cols.append(runs)
    for c in cols:
        out = apply_network(ces, c.astype(np.int64))
        assert (np.diff(out, axis=0) >= 0).all()
        assert np.array_equal(np.sort(c, axis=0), out)


def mid_kernel_schedule(col, n, npad, net64):
    """ssc_stack_mid.cu steps 3, 4a, 4b for one column: runs of 64 through the register network, bitonic
    merges (first stage pairs k with L-1-k of the next run, half-cleaners down to 64, then the last six
    stages of every 64-block) inside each half of the column, then the shared last level."""
    v = np.full(npad, 0xFFFF, dtype=np.int64)
    v[:n] = col
    half = npad // 2

    def block_merges(first, length):
        for c0 in range(first, first + length, 64):
            d = 32
            while d >= 1:
                for i in range(64):
                    if (i & d) == 0:
                        a, b = c0 + i, c0 + i + d
                        v[a], v[b] = min(v[a], v[b]), max(v[a], v[b])
                d >>= 1

    def half_cleaners(first, length, d_first):
        d = d_first
        while d >= 64:
            for j in range(length // 2):
                blk, k = j // d, j % d
                a = first + blk * 2 * d + k
                b = a + d
                v[a], v[b] = min(v[a], v[b]), max(v[a], v[b])
            d >>= 1

    for h0 in (0, half):
        n_end = min(max(n, h0), h0 + half)
        for c0 in range(h0, n_end, 64):
            blk = v[c0:c0 + 64].copy()
            for a, b in net64:
                blk[a], blk[b] = min(blk[a], blk[b]), max(blk[a], blk[b])
            v[c0:c0 + 64] = blk
        L = 64
        while L < half:
            for j in range(half // 2):
                blk, k = j // L, j % L
                a, b = h0 + blk * 2 * L + k, h0 + blk * 2 * L + 2 * L - 1 - k
                v[a], v[b] = min(v[a], v[b]), max(v[a], v[b])
            half_cleaners(h0, half, L >> 1)
            block_merges(h0, half)
            L <<= 1
    for j in range(half):
        a, b = j, npad - 1 - j
        v[a], v[b] = min(v[a], v[b]), max(v[a], v[b])
    for h0 in (0, half):
        half_cleaners(h0, half, half >> 1)
        block_merges(h0, half)
    return v


@pytest.mark.parametrize("n,npad", [(65, 128), (100, 128), (128, 128), (129, 256), (200, 256), (256, 256),
                                    (300, 512), (512, 512), (700, 1024), (1024, 1024)])
def test_mid_kernel_merge_schedule_sorts(n, npad):
    net64 = networks()[64]
    rng = np.random.default_rng(n)
    for trial in range(3):
        col = rng.integers(0, 65536 if trial else 4, size=n)
        out = mid_kernel_schedule(col, n, npad, net64)
        assert np.array_equal(out[:n], np.sort(col))
        assert (out[n:] == 0xFFFF).all()


def mid_kernel_schedule_virtual(col, n, npad, net64):
    """The same schedule as ssc_stack_mid.cu runs it since the virtual sentinels: only n_alloc = n rounded up to 8 rows are
    STORED (an index at or beyond n_alloc raises here), every loop is bounded as in the kernel -- a compare-exchange whose
    upper element would be a sentinel at or beyond n_alloc is skipped -- and one loop over the levels L = 32 (the runs of
    64 through the network), 64 .. HALF/2 (merges inside a half), HALF (the shared last level) drives one register step."""
    n_alloc = min(npad, (n + 7) & ~7)
    v = np.full(n_alloc, 0xFFFF, dtype=np.int64)        # rows n .. n_alloc-1: stored sentinels
    v[:n] = col
    half = npad // 2

    def ce(a, b):
        assert a < b < n_alloc
        v[a], v[b] = min(v[a], v[b]), max(v[a], v[b])

    def register_step(c0, sort):                        # load_64 / sort_network or bitonic_merge_64 / store_64
        blk = np.array([v[c0 + k] if c0 + k < n_alloc else 0xFFFF for k in range(64)], dtype=np.int64)
        if sort:
            for a, b in net64:
                blk[a], blk[b] = min(blk[a], blk[b]), max(blk[a], blk[b])
        else:
            d = 32
            while d >= 1:
                for i in range(64):
                    if (i & d) == 0:
                        blk[i], blk[i + d] = min(blk[i], blk[i + d]), max(blk[i], blk[i + d])
                d >>= 1
        for k in range(64):
            if c0 + k < n_alloc:
                v[c0 + k] = blk[k]
            else:
                assert blk[k] == 0xFFFF                 # what is not stored must still be a sentinel

    def half_cleaners(first, length, d_first):
        d = d_first
        while d >= 64:
            for sb in range(first, first + length, 2 * d):
                for i in range(sb, min(sb + d, n_alloc - d)):
                    ce(i, i + d)
            d >>= 1

    L = 32
    while L <= half:                                    # (the two threads of a pair run this loop side by side)
        for hf in (0, 1):
            h0 = hf * half
            n_end = min(max(n, h0), h0 + half)
            if 64 <= L < half:
                for base in range(h0, h0 + half, 2 * L):
                    for k in range(max(0, base + 2 * L - n_alloc), L):
                        ce(base + k, base + 2 * L - 1 - k)
            elif L == half:
                for j in range(max(hf * (half // 2), npad - n_alloc), (hf + 1) * (half // 2)):
                    ce(j, npad - 1 - j)
        for hf in (0, 1):                               # (after the barrier that follows the shared stage)
            h0 = hf * half
            n_end = min(max(n, h0), h0 + half)
            if L >= 64:
                half_cleaners(h0, half, L >> 1)
            c_end = n_end if L == 32 else min(h0 + half, n_alloc)
            for c0 in range(h0, c_end, 64):
                register_step(c0, L == 32)
        L <<= 1
    return v, n_alloc


@pytest.mark.parametrize("n,npad", [(65, 128), (72, 128), (100, 128), (127, 128), (128, 128), (129, 256), (136, 256), (193, 256),
                                    (200, 256), (255, 256), (256, 256), (257, 512), (300, 512), (449, 512), (512, 512),
                                    (513, 1024), (700, 1024), (1000, 1024), (1024, 1024)])
def test_mid_kernel_schedule_with_virtual_sentinels_sorts(n, npad):
    net64 = networks()[64]
    rng = np.random.default_rng(1000 + n)
    for trial in range(4):
        hi = (4, 65536, 65536, 2)[trial]
        col = rng.integers(0, hi, size=n)
        if trial == 2:
            col[rng.integers(0, n, size=n // 3)] = 65535        # real samples equal to the sentinel value
        if trial == 3:
            col = col * 65535
        out, n_alloc = mid_kernel_schedule_virtual(col, n, npad, net64)
        assert np.array_equal(out[:n], np.sort(col))
        assert (out[n:n_alloc] == 0xFFFF).all()
        assert np.array_equal(out[:n], mid_kernel_schedule(col, n, npad, net64)[:n])


def test_single_precision_no_rejection_proof_is_conservative():
    """ssc_stack_fast.cu:clip_cannot_fire decides "the first sigma-clipping pass rejects nothing" in FP32;
    it must never say so when the exact FP64 test (M2 - 2 min > floor(2 sig_low sigma), likewise above)
    would reject.  Restated in numpy on random and on adversarial columns (extremes placed right at the
    threshold); the GPU's approximate square root is covered by widening the FP32 estimate by 4 ulp."""
    rng = np.random.default_rng(5)
    n = 64
    sig_low, sig_high = 4.0, 3.0
    cols = []
    base = rng.normal(3000, 40, size=(200000, n))
    cols.append(base)
    for k in (2.4, 2.9, 3.0, 3.05, 3.9, 4.0, 4.05):                 # push one sample to about k sigma
        c = rng.normal(20000, 300, size=(60000, n))
        c[:, 0] = 20000 + k * 300 * rng.normal(1.0, 0.01, size=60000)
        cols.append(c)
        c = c.copy()
        c[:, 0] = 20000 - k * 300 * rng.normal(1.0, 0.01, size=60000)
        cols.append(c)
    cols.append(np.full((10, n), 65535.0))
    cols.append(rng.integers(0, 2, size=(1000, n)).astype(float) * 65535)
    x = np.clip(np.rint(np.concatenate(cols)), 0, 65535).astype(np.int64)
    x.sort(axis=1)
    S, Q = x.sum(1), (x * x).sum(1)
    D = n * Q - S * S
    M2 = x[:, n // 2 - 1] + x[:, n // 2]
    d_lo, d_hi = M2 - 2 * x[:, 0], 2 * x[:, -1] - M2
    # exact path (ssc_device.cuh:sd_from_sums + the integer thresholds of the queue pass)
    sigma = np.sqrt((D.astype(np.float64) / float(n * n)) * (float(n) / float(n - 1)))
    TL = np.floor(np.minimum(2.0 * (sig_low * sigma), 1073741824.0))
    TH = np.floor(np.minimum(2.0 * (sig_high * sigma), 1073741824.0))
    rejects = (d_lo > TL) | (d_hi > TH)
    # FP32 proof, with the square root 4 ulp too LARGE (the worst direction for the proof)
    f32 = np.float32
    s32 = np.sqrt(D.astype(f32) * f32(1.0 / (n * (n - 1)))).astype(f32) * f32(1.0 + 4 * 2.0 ** -23)
    t_lo = f32(2.0) * f32(sig_low) * s32 * f32(0.9999) - f32(1.0)
    t_hi = f32(2.0) * f32(sig_high) * s32 * f32(0.9999) - f32(1.0)
    clear = ((d_lo == 0) | (d_lo.astype(f32) < t_lo)) & ((d_hi == 0) | (d_hi.astype(f32) < t_hi))
    assert not (clear & rejects).any()
    assert clear.mean() > 0.3 and rejects.mean() > 0.1              # the sweep exercises both outcomes
