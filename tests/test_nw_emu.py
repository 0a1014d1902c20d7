"""CPU tests of the warp-wavefront NW (kma_b200/csrc/kmagpu_nw.cuh) through the lock-step emulator in
tests/emu/nw_emu.cpp: the per-lane step, geometry, start-cell and walk functions are the source the CUDA kernel
compiles. Checked against the oracle's NW_score / NW_band_score restatement (pinned to the reference by
test_oracle_align.py) for every k mode, full and banded, ragged sizes and N bases."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libnwemu.so")
    src = os.path.join(HERE, "emu", "nw_emu.cpp")
    hdr = os.path.join(util.ROOT, "kma_b200", "csrc", "kmagpu_nw.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def pack(seq):
    """2-bit pack MSB first (compdna.h), padded with two zero words"""
    n = len(seq)
    words = np.zeros(n // 32 + 3, dtype=np.uint64)
    for i, b in enumerate(seq):
        words[i >> 5] |= np.uint64(int(b) & 3) << np.uint64(62 - 2 * (i & 31))
    return words


def problem(rng, t_len, q_len, err=0.08, n_rate=0.01):
    """query = template-like sequence with substitutions / indels so that the optimal path is non-trivial"""
    t = rng.integers(0, 4, size=t_len + 64).astype(np.uint8)
    q = []
    i = 0
    while len(q) < q_len:
        r = rng.random()
        if r < err / 3:
            i += 1
        elif r < 2 * err / 3:
            q.append(rng.integers(0, 4))
        elif r < err:
            q.append((t[i % len(t)] + 1 + rng.integers(0, 3)) % 4)
            i += 1
        else:
            q.append(t[i % len(t)])
            i += 1
    q = np.array(q[:q_len], dtype=np.uint8)
    q[rng.random(q_len) < n_rate] = 4
    return t, q


def run_both(emu, t, q, k, t_s, t_e, q_s, q_e, band):
    L = util.orc()
    pen = util.oracle_params()
    tw = pack(t)
    want = (C.c_int * 6)()
    L.orc_nw(pen, tw.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), k, t_s, t_e, q_s, q_e, band, want)
    # emulator penalties: W1 U MM M d[25]
    p = list(pen)
    pen29 = (C.c_int * 29)(p[3], p[2], p[1], p[0], *p[7:32])
    res = []
    t_len, q_len = t_e - t_s, q_e - q_s
    maps = []
    # rs = 1: the row sweep for rows of up to 256 cells (what the kernel runs); rs = 0: the wavefront for every size
    for order, d8, rs in ((0, 1, 1), (0, 1, 0), (1, 1, 0), (0, 0, 0)):
        got = (C.c_int * 6)()
        emap = np.zeros(max(1, t_len * q_len), dtype=np.uint8)
        rc = emu.emu_nw2(pen29, tw.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), k, t_s, t_e, q_s, q_e, band,
                         order, d8, rs, got, None, emap.ctypes.data_as(C.c_void_p))
        assert rc == 0, (rc, k, t_len, q_len, band, rs)
        res.append(list(got))
        maps.append(emap)
    assert res[1] == res[2] == res[3], "lane order / table form changes the result: same-step hazard"
    assert res[0] == res[1], ("row sweep != wavefront", k, t_len, q_len, band, res[0], res[1])
    if t_len and q_len:
        assert np.array_equal(maps[0], maps[1]), ("traceback bytes of the row sweep differ", k, t_len, q_len, band,
                                                  np.argwhere(maps[0] != maps[1])[:5])
    return list(want), res[0]


@pytest.mark.parametrize("k", [0, -1, -2, 1, 2])
def test_full_matrix(emu, k):
    rng = np.random.default_rng(100 + k)
    sizes = [(1, 1), (1, 2), (2, 1), (3, 7), (31, 31), (32, 32), (33, 33), (32, 5), (5, 32), (64, 40), (65, 33), (100, 20),
             (20, 100), (97, 130), (150, 214), (40, 300), (300, 40), (129, 3), (3, 129)]
    sizes += [(int(rng.integers(1, 200)), int(rng.integers(1, 200))) for _ in range(60)]
    for t_len, q_len in sizes:
        t, q = problem(rng, t_len, q_len + 8, err=rng.choice([0.0, 0.05, 0.2, 0.6]))
        t_s, q_s = int(rng.integers(0, 40)), int(rng.integers(0, 8))
        want, got = run_both(emu, t, q, k, t_s, t_s + t_len, q_s, q_s + q_len, 0)
        assert got == want, (k, t_len, q_len, want, got)


@pytest.mark.parametrize("k", [0, -1, -2, 1, 2])
def test_banded(emu, k):
    rng = np.random.default_rng(200 + k)
    cases = []
    for _ in range(50):
        t_len = int(rng.integers(70, 700))
        q_len = max(66, t_len + int(rng.integers(-60, 60)))
        band = abs(t_len - q_len) + 64
        if q_len <= band or t_len <= band:
            continue
        cases.append((t_len, q_len, band))
    cases += [(200, 200, 64), (201, 200, 65), (130, 129, 65), (500, 431, 133), (431, 500, 133), (1000, 1000, 64),
              (300, 300, 100), (300, 290, 150), (1200, 900, 364), (700, 700, 300)]
    assert len(cases) > 30
    for t_len, q_len, band in cases:
        t, q = problem(rng, t_len, q_len, err=rng.choice([0.02, 0.1, 0.3]))
        t_s = int(rng.integers(0, 40))
        want, got = run_both(emu, t, q, k, t_s, t_s + t_len, 0, q_len, band)
        assert got == want, (k, t_len, q_len, band, want, got)


def test_trivial_and_unrelated(emu):
    rng = np.random.default_rng(7)
    t, q = problem(rng, 50, 50)
    for k in (0, -1, 1):
        assert run_both(emu, t, q, k, 3, 3, 0, 10, 0)[0] == run_both(emu, t, q, k, 3, 3, 0, 10, 0)[1]
        want, got = run_both(emu, t, q, k, 3, 13, 5, 5, 0)
        assert want == got
    # unrelated sequences: gap-heavy paths, many ties
    for _ in range(30):
        t = rng.integers(0, 4, size=200).astype(np.uint8)
        q = rng.integers(0, 4, size=200).astype(np.uint8)
        t_len, q_len, k = int(rng.integers(1, 120)), int(rng.integers(1, 120)), int(rng.choice([0, -1, -2, 1, 2]))
        want, got = run_both(emu, t, q, k, 0, t_len, 0, q_len, 0)
        assert want == got, (k, t_len, q_len)
    # low-complexity: maximal number of score ties
    for _ in range(30):
        t = np.zeros(300, dtype=np.uint8); q = np.zeros(300, dtype=np.uint8)
        t[rng.integers(0, 300, size=10)] = 1
        q[rng.integers(0, 300, size=10)] = 1
        t_len, q_len, k = int(rng.integers(1, 150)), int(rng.integers(1, 150)), int(rng.choice([0, -1, -2, 1, 2]))
        want, got = run_both(emu, t, q, k, 0, t_len, 0, q_len, 0)
        assert want == got, (k, t_len, q_len)


@pytest.mark.parametrize("k", [0, -1, -2, 1, 2])
def test_thread_per_problem(emu, k):
    """nw_thread (the one-thread-per-problem NW of the queue kernels) against the oracle and, byte for byte of the
    traceback matrix, against the wavefront"""
    rng = np.random.default_rng(300 + k)
    L = util.orc()
    pen = util.oracle_params()
    p = list(pen)
    pen29 = (C.c_int * 29)(p[3], p[2], p[1], p[0], *p[7:32])
    sizes = [(1, 1), (1, 2), (2, 1), (3, 7), (8, 16), (16, 8), (31, 31), (64, 32), (128, 64), (5, 64), (100, 3), (1, 64)]
    sizes += [(int(rng.integers(1, 129)), int(rng.integers(1, 65))) for _ in range(120)]
    for t_len, q_len in sizes:
        kind = rng.integers(0, 3)
        if kind == 0:
            t, q = problem(rng, t_len, q_len + 8, err=rng.choice([0.0, 0.05, 0.2, 0.6]))
        elif kind == 1:
            t = rng.integers(0, 4, size=t_len + 64).astype(np.uint8); q = rng.integers(0, 4, size=q_len + 8).astype(np.uint8)
        else:
            t = np.zeros(t_len + 64, dtype=np.uint8); q = np.zeros(q_len + 8, dtype=np.uint8)
            t[rng.integers(0, t_len + 64, size=6)] = 1; q[rng.integers(0, q_len + 8, size=4)] = 1
        t_s, q_s = int(rng.integers(0, 40)), int(rng.integers(0, 8))
        tw = pack(t)
        want = (C.c_int * 6)()
        L.orc_nw(pen, tw.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), k, t_s, t_s + t_len, q_s, q_s + q_len, 0, want)
        ref_map = np.zeros(t_len * q_len, dtype=np.uint8)
        got0 = (C.c_int * 6)()
        assert emu.emu_nw2(pen29, tw.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), k, t_s, t_s + t_len, q_s, q_s + q_len, 0,
                           0, 1, 0, got0, None, ref_map.ctypes.data_as(C.c_void_p)) == 0
        for d8, stride in ((1, 1), (0, 1), (1, 32)):
            got = (C.c_int * 6)()
            emap = np.zeros(t_len * q_len, dtype=np.uint8)
            assert emu.emu_nw_thread(pen29, tw.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), k, t_s, t_s + t_len, q_s,
                                     q_s + q_len, d8, stride, got, emap.ctypes.data_as(C.c_void_p)) == 0
            assert list(got) == list(want), (k, t_len, q_len, d8, stride, list(want), list(got))
            assert np.array_equal(emap, ref_map), (k, t_len, q_len)
