"""CPU-side checks of the product's host logic: the C-ABI library loads and exports every symbol the
header declares, the cell counter matches the oracle's loop-bound definition, and the kernel's
per-lane recurrence / band logic (k1_core.cuh, compiled for the host) reproduces the reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, golden_names
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode


def test_library_loads_and_exports_declared_symbols():
    L = P.load_library()
    names = P.declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), "libprrn_gpu.so does not export " + n
    assert b"sm_100a" in L.pg_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(P.PgError) as e:
        P.Context(0)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)


def test_cell_count_matches_oracle(oracle):
    rng = np.random.default_rng(5)
    lens = [0, 1, 2, 7, 33, 100, 257, 400, 513]
    enc = [rng.integers(3, 23, size=n).astype(np.uint8) for n in lens]
    ss = P.SeqSet(enc)
    for sh in (-60, -50, -100, -5, 0, 3, 50, 1000):
        prm = P.Params(P.ALPRM(sh=sh))
        want = 0
        for j in range(1, len(enc)):
            for i in range(j):
                want += oracle.band_cells(oracle.seq(enc[i]), oracle.seq(enc[j]), sh)
        assert P.calcdist_cells(ss, prm) == want, sh
    # a sub-range of the condensed index
    prm = P.Params(P.ALPRM(sh=-60))
    n = len(enc)
    tot = sum(P.calcdist_cells(ss, prm, k, k + 1) for k in range(n * (n - 1) // 2))
    assert tot == P.calcdist_cells(ss, prm)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(ROOT, "tests", "host_emul", "k1_emul.cc")
    out = os.path.join(ROOT, "tests", "host_emul", "libk1emul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    L = C.CDLL(out)
    L.k1_emul_score.restype = C.c_int
    return L


def _emul_score(L, q, s, mtx, u, v, sh, R):
    q = np.ascontiguousarray(q, np.uint8)
    s = np.ascontiguousarray(s, np.uint8)
    m = np.ascontiguousarray(mtx, np.int32)
    return L.k1_emul_score(q.ctypes.data_as(C.c_void_p), len(q), s.ctypes.data_as(C.c_void_p), len(s),
                           m.ctypes.data_as(C.c_void_p), m.shape[0], u, v, sh, -v, -u, -v, -u, R)


@pytest.mark.parametrize("name", ["score_p24_blosum62", "score_p24_sh3", "score_p24_sh0", "score_p24_u3v11",
                                  "score_ragged", "score_long1300", "score_c1_ce13a"])
def test_kernel_recurrence_emulation_matches_reference(emul, name):
    """k1_core.cuh (the code the CUDA kernel runs per lane), emulated warp-wide on the host."""
    g = golden(name)
    enc = [seqcode.encode_protein(x) for x in g["seqs"]]
    M = np.array(g["matrix"]).astype(np.int32)
    u, v, sh = int(float(g["params"]["u"])), int(float(g["params"]["v"])), int(g["params"]["sh"])
    n = len(enc)
    for R in (4, 16):
        for j in range(1, n):
            for i in range(j):
                want = g["scores"][j * (j - 1) // 2 + i]
                assert _emul_score(emul, enc[i], enc[j], M, u, v, sh, R) == want
                assert _emul_score(emul, enc[j], enc[i], M, u, v, sh, R) == want  # rows/cols swapped


def test_kernel_recurrence_fuzz_small_adversarial(emul, oracle):
    """Tiny sequences, tiny alphabets, narrow bands, v in {0..12}: the corner cases where band cuts
    and boundary openings decide the score (found a top-boundary vertical-open bug in round 1)."""
    M = np.array(golden("score_p24_blosum62")["matrix"])
    Mi = np.nan_to_num(M).astype(np.int32)
    rng = np.random.default_rng(2024)
    for _ in range(4000):
        la, lb = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        hi = 3 + int(rng.choice([2, 4, 20]))
        a = rng.integers(3, hi, size=la).astype(np.uint8)
        b = rng.integers(3, hi, size=lb).astype(np.uint8)
        sh = int(rng.choice([-100, -60, -30, -10, 0, 1, 2, 3, 5, 100]))
        u, v = int(rng.choice([1, 2, 3])), int(rng.choice([0, 1, 5, 9, 12]))
        want = oracle.aln_score_d(oracle.seq(a), oracle.seq(b), M, oracle.params(u=u, v=v, sh=sh))
        R = int(rng.choice([4, 8, 16]))
        got = emul.k1_emul_score(a.ctypes.data_as(C.c_void_p), la, b.ctypes.data_as(C.c_void_p), lb,
                                 Mi.ctypes.data_as(C.c_void_p), 25, u, v, sh, -v, -u, -v, -u, R)
        assert got == want, (la, lb, sh, u, v, R)


@pytest.fixture(scope="module")
def emul2():
    src = os.path.join(ROOT, "tests", "host_emul", "k2_emul.cc")
    out = os.path.join(ROOT, "tests", "host_emul", "libk2emul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    L = C.CDLL(out)
    L.k2_emul_align.restype = C.c_int
    return L


def _emul_align(L, q, s, Mi, u, v, sh, R):
    q = np.ascontiguousarray(q, np.uint8)
    s = np.ascontiguousarray(s, np.uint8)
    out = np.zeros(2 * (len(q) + len(s) + 8), np.int32)
    sc = C.c_int(0)
    n = L.k2_emul_align(q.ctypes.data_as(C.c_void_p), len(q), s.ctypes.data_as(C.c_void_p), len(s),
                        Mi.ctypes.data_as(C.c_void_p), Mi.shape[0], u, v, sh, R, out.ctypes.data_as(C.c_void_p),
                        C.byref(sc))
    return sc.value, [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]


@pytest.mark.parametrize("name", ["align_p16_blosum62", "align_p16_sh3", "align_p16_u3v11", "align_ragged",
                                  "align_long1300", "align_c1_ce13a"])
def test_alignment_kernel_emulation_matches_reference(emul2, name):
    """k2_core.cuh (direction bits, traceback, record replay) + host stdskl against align2 goldens."""
    g = golden(name)
    enc = [seqcode.encode_protein(x) for x in g["seqs"]]
    Mi = np.nan_to_num(np.array(g["matrix"])).astype(np.int32)
    u, v, sh = int(float(g["params"]["u"])), int(float(g["params"]["v"])), int(g["params"]["sh"])
    for p in g["pairs"]:
        for R in (4, 16):
            sc, raw = _emul_align(emul2, enc[p["i"]], enc[p["j"]], Mi, u, v, sh, R)
            assert sc == p["score"]
            assert P.stdskl(raw) == [tuple(x) for x in p["skl"]]


def test_alignment_kernel_emulation_fuzz(emul2, oracle):
    M = np.array(golden("score_p24_blosum62")["matrix"])
    Mi = np.nan_to_num(M).astype(np.int32)
    rng = np.random.default_rng(31)
    for _ in range(3000):
        la, lb = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        hi = 3 + int(rng.choice([2, 4, 20]))
        a = rng.integers(3, hi, size=la).astype(np.uint8)
        b = rng.integers(3, hi, size=lb).astype(np.uint8)
        sh = int(rng.choice([-100, -60, -30, -10, 0, 1, 2, 3, 5, 100]))
        u, v = int(rng.choice([1, 2, 3])), int(rng.choice([0, 1, 5, 9, 12]))
        osc, olist = oracle.align_ngp(oracle.seq(a), oracle.seq(b), M, oracle.params(u=u, v=v, sh=sh), std=False)
        sc, raw = _emul_align(emul2, a, b, Mi, u, v, sh, int(rng.choice([4, 8, 16])))
        assert sc == osc and raw == olist, (la, lb, sh, u, v)


@pytest.fixture(scope="module")
def emulp():
    src = os.path.join(ROOT, "tests", "host_emul", "k1p_emul.cc")
    out = os.path.join(ROOT, "tests", "host_emul", "libk1pemul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


def test_packed_int16x2_recurrence_emulation(emulp, oracle):
    """k1p_core.cuh: two alignments per register with int16 wrap-around arithmetic emulated on the
    host; small adversarial cases + realistic lengths incl. multi-pass and unequal query lengths."""
    M = np.array(golden("score_p24_blosum62")["matrix"])
    Mi = np.nan_to_num(M).astype(np.int32)

    def run(q0, q1, s, u, v, sh, R):
        o0, o1 = C.c_int(0), C.c_int(0)
        emulp.k1p_emul_score2(q0.ctypes.data_as(C.c_void_p), len(q0), q1.ctypes.data_as(C.c_void_p), len(q1),
                              s.ctypes.data_as(C.c_void_p), len(s), Mi.ctypes.data_as(C.c_void_p), 25, u, v, sh, R,
                              C.byref(o0), C.byref(o1))
        return o0.value, o1.value

    rng = np.random.default_rng(8)
    for _ in range(2500):
        hi = 3 + int(rng.choice([2, 4, 20]))
        q0, q1, s = (rng.integers(3, hi, size=int(rng.integers(1, 40))).astype(np.uint8) for _ in range(3))
        sh = int(rng.choice([-100, -60, -30, -10, 0, 1, 2, 3, 5, 100]))
        u, v = int(rng.choice([1, 2, 3])), int(rng.choice([0, 1, 5, 9, 12]))
        p = oracle.params(u=u, v=v, sh=sh)
        want = (oracle.aln_score_d(oracle.seq(q0), oracle.seq(s), M, p), oracle.aln_score_d(oracle.seq(q1), oracle.seq(s), M, p))
        assert run(q0, q1, s, u, v, sh, int(rng.choice([4, 8, 16]))) == want
    g = golden("score_long1300")
    enc = [seqcode.encode_protein(x) for x in g["seqs"]]
    enc += [enc[0][:300], enc[1][:513]]
    p = oracle.params()
    for a in range(0, len(enc) - 1, 2):
        for b in range(len(enc)):
            want = (oracle.aln_score_d(oracle.seq(enc[a]), oracle.seq(enc[b]), M, p),
                    oracle.aln_score_d(oracle.seq(enc[a + 1]), oracle.seq(enc[b]), M, p))
            assert run(enc[a], enc[a + 1], enc[b], 2, 9, -60, 16) == want
    # the 16-bit range proof obligation
    assert emulp.k1p_emul_fits(15, 0, 9, 1300) == 1
    assert emulp.k1p_emul_fits(15, 0, 9, 2200) == 0
    assert emulp.k1p_emul_fits(13, -2, 9, 1300) == 1


def test_packed_plan_covers_every_pair_exactly_once():
    """The tournament schedule of the packed kernel (query pairs x subjects with two valid bits):
    every condensed index of the requested range exactly once, for full triangles and shards."""
    rng = np.random.default_rng(12)
    for n in (2, 3, 4, 5, 8, 9, 64, 301):
        enc = [rng.integers(3, 23, size=int(rng.integers(1, 600))).astype(np.uint8) for _ in range(n)]
        ss = P.SeqSet(enc)
        npair = n * (n - 1) // 2
        ranges = [(0, npair)] + [(npair * r // w, npair * (r + 1) // w) for w in (2, 3, 8) for r in range(w)]
        ranges += [(1, 1), (npair - 1, npair)]
        for k0, k1 in ranges:
            items, slots, cover = P.packed_plan_coverage(ss, k0, k1)
            assert np.all(cover == 1), (n, k0, k1)
            if k1 - k0 > 1000:
                assert 2 * slots <= 1.05 * (k1 - k0) + 2 * n     # halves are almost never idle
