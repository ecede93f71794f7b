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
    for m in re.finditer(r"#define SSC_SORTNET_(\d+)((?:.*\\\n)*.*\n)", text):
        nets[int(m.group(1))] = [(int(a), int(b)) for a, b in re.findall(r"SSC_CE\((\d+),(\d+)\)", m.group(2))]
    return nets


def apply_network(ces, v):
    v = v.copy()
    for a, b in ces:                     # SSC_CE(a, b): min to v[a], max to v[b]; vectorised over the columns of v
        lo, hi = np.minimum(v[a], v[b]), np.maximum(v[a], v[b])
        v[a], v[b] = lo, hi
    return v


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
