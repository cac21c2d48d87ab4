"""k1f_core.cuh (the code kernel K1F runs per lane: floating-point recurrence, boundary tables, band
cut, lastD scans, SWG clamp, Fwd2d_vd records) emulated warp-wide on the host against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "host_emul", "k1f_emul.cc")
    out = os.path.join(ROOT, "tests", "host_emul", "libk1femul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    L = C.CDLL(out)
    L.k1f_emul_score.restype = C.c_double
    L.k1f_emul_score.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                 C.c_double, C.c_double, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]
    return L


def _flags(x, l, r, ex):
    return (ex & 3) | (4 if l else 0) | (8 if r != len(x) else 0)


def _run(L, M, a, al, ar, aex, b, bl, br, bex, u, v, tg, sh, R, vt, swap, mode):
    if not swap:
        q, ql, qr, qf, s, sl, sr, sf, mm = a, al, ar, _flags(a, al, ar, aex), b, bl, br, _flags(b, bl, br, bex), M
    else:
        q, ql, qr, qf, s, sl, sr, sf, mm = b, bl, br, _flags(b, bl, br, bex), a, al, ar, _flags(a, al, ar, aex), np.ascontiguousarray(M.T)
    qq = np.ascontiguousarray(q[ql:qr]) if qr > ql else np.zeros(1, np.uint8)
    ss = np.ascontiguousarray(s[sl:sr]) if sr > sl else np.zeros(1, np.uint8)
    ends = (C.c_int * 2)(0, 0)
    r = L.k1f_emul_score(qq.ctypes.data, qr - ql, qf, ss.ctypes.data, sr - sl, sf, mm.ctypes.data, M.shape[0], u, v, tg,
                         sh, R, vt, mode, swap, bl - al, ends)
    return r, (ends[0], ends[1])


def test_float_recurrence_all_modes_fuzz(emul, oracle):
    M = np.nan_to_num(np.array(golden("score_p24_pam_f32")["matrix"], dtype=np.float64))
    rng = np.random.default_rng(5)
    for it in range(2500):
        la, lb = int(rng.integers(0, 45)), int(rng.integers(0, 45))
        if it % 60 == 0:
            la, lb = int(rng.integers(100, 700)), int(rng.integers(100, 700))
        hi = 3 + int(rng.choice([2, 4, 20]))
        a = rng.integers(3, hi, size=la).astype(np.uint8)
        b = rng.integers(3, hi, size=lb).astype(np.uint8)
        sh = int(rng.choice([-100, -60, -30, -10, 0, 1, 2, 3, 5, 100]))
        u, v = float(rng.choice([1, 2, 3, 0.6, 1.5])), float(rng.choice([0, 1, 5, 9, 12, 4.5]))
        tg = float(rng.choice([1, 1, 0.5, 0, 0.3]))
        vt = int(rng.integers(0, 2))
        al = int(rng.integers(0, la + 1)) if rng.random() < 0.3 else 0
        ar = int(rng.integers(al, la + 1)) if rng.random() < 0.3 else la
        bl = int(rng.integers(0, lb + 1)) if rng.random() < 0.3 else 0
        br = int(rng.integers(bl, lb + 1)) if rng.random() < 0.3 else lb
        aex, bex = int(rng.choice([0, 0, 1, 2, 3])), int(rng.choice([0, 0, 1, 2, 3]))
        mode = int(rng.choice([0, 1, 2]))
        p = oracle.params(u=u, v=v, sh=sh, tgapf=tg, vtype=vt, lcl=16 if mode == 1 else 0)
        want, we = oracle.aln_score_full(oracle.seq(a, al, ar, aex & 1, (aex >> 1) & 1),
                                         oracle.seq(b, bl, br, bex & 1, (bex >> 1) & 1), M, p, want_ends=(mode == 2))
        for swap in ((0,) if mode == 2 else (0, 1)):        # Fwd2d_vd is not symmetric: reference orientation only
            got, ge = _run(emul, M, a, al, ar, aex, b, bl, br, bex, u, v, tg, sh, int(rng.choice([4, 8, 16])), vt, swap, mode)
            assert got == want, (mode, la, lb, sh, u, v, tg, vt, swap)
            if mode == 2:
                assert ge == we, (la, lb, sh, u, v, tg, vt)
