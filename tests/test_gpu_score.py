"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors frozen from the
unmodified reference, the oracle on seeded inputs, and size-independent properties at full size."""
import numpy as np
import pytest

from conftest import golden, golden_names
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode

pytestmark = pytest.mark.gpu

import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_synth  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = P.Context(0)
    yield c
    c.close()


def _params(g):
    h = g["params"]
    a = P.ALPRM(u=float(h["u"]), v=float(h["v"]), tgapf=float(h["tgapf"]), scale=float(h["scale"]),
                u1=float(h["u1"]), k1=int(h["k1"]), ls=int(h["ls"]), sh=int(h["sh"]))
    return P.Params(a, lcl=int(h["lcl"]), vtype=1 if h["vtype"] == "f64" else 0)


def _integer_case(g):
    # codes 3..22 are the 20 standard residues of the synthetic sets; the reference leaves a few
    # entries of rarely used codes (e.g. mtx[SEC][SEC]) uninitialised, so only look at those rows
    M = np.array(g["matrix"])[3:23, 3:23]
    # lcl variants / tgapf < 1 / non-integral matrices run on kernel K1F: tests/test_gpu_scoref.py
    return float(g["params"]["tgapf"]) == 1.0 and int(g["params"]["lcl"]) == 0 and np.all(M == np.rint(M))


INT_GOLDENS = [n for n in golden_names("score_") if _integer_case(golden(n))]


@pytest.mark.parametrize("name", INT_GOLDENS)
def test_golden_scores_and_dist_bit_exact(ctx, name):
    g = golden(name)
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    ss = P.SeqSet(enc)
    prm = _params(g)
    M = np.array(g["matrix"])
    n = len(enc)
    dist = ctx.calcdist(ss, prm, M)
    assert dist.dtype == prm.ftype
    assert np.array_equal(dist.astype(np.float64), np.array(g["dist"])), "calcdist differs from the reference"
    ia = [i for j in range(1, n) for i in range(j)]
    ib = [j for j in range(1, n) for i in range(j)]
    want = np.array(g["scores"])
    assert np.array_equal(ctx.score_pairs(ss, ia, ib, prm, M).astype(np.float64), want)
    assert np.array_equal(ctx.score_pairs(ss, ib, ia, prm, M).astype(np.float64), want)  # swapped roles


def test_double_vtype_matches_float_golden(ctx):
    g = golden("score_p24_blosum62")
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    prm = _params(g)
    prm.vtype = 1
    n = len(enc)
    ia = [i for j in range(1, n) for i in range(j)]
    ib = [j for j in range(1, n) for i in range(j)]
    sc = ctx.score_pairs(P.SeqSet(enc), ia, ib, prm, np.array(g["matrix"]))
    assert sc.dtype == np.float64 and np.array_equal(sc, np.array(g["scores"]))


def test_against_oracle_c2_subset(ctx, oracle):
    seqs = gen_synth.config_set("c2", 90)
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = np.array(golden("score_p24_blosum62")["matrix"])
    for sh, vt in ((-60, 0), (-60, 1), (-25, 0), (40, 0)):
        prm = P.Params(P.ALPRM(sh=sh), vtype=vt)
        op = oracle.params(sh=sh, vtype=vt)
        want, raw = oracle.calcdist([oracle.seq(e) for e in enc], M, op)
        got = ctx.calcdist(P.SeqSet(enc), prm, M)
        assert np.array_equal(got.astype(np.float64), want), (sh, vt)


def test_int32_and_packed_kernels_agree(ctx, oracle, monkeypatch):
    """calcdist runs the packed int16x2 kernel when values provably fit 16 bits; PG_FORCE_INT32=1
    routes the same call through the int32 kernel.  Both must equal the oracle."""
    seqs = gen_synth.config_set("c2", 70) + gen_synth.synth_set(5, 900, 0.1, 0.5, 9)
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = np.array(golden("score_p24_blosum62")["matrix"])
    ss = P.SeqSet(enc)
    for sh in (-60, 7):
        prm = P.Params(P.ALPRM(sh=sh), vtype=1)
        want, _ = oracle.calcdist([oracle.seq(e) for e in enc], M, oracle.params(sh=sh, vtype=1))
        monkeypatch.delenv("PG_FORCE_INT32", raising=False)
        packed = ctx.calcdist(ss, prm, M)
        monkeypatch.setenv("PG_FORCE_INT32", "1")
        plain = ctx.calcdist(ss, prm, M)
        monkeypatch.delenv("PG_FORCE_INT32", raising=False)
        assert np.array_equal(packed, want), sh
        assert np.array_equal(plain, want), sh


def test_packed_kernel_every_rows_variant(ctx, oracle):
    """Mixed lengths 20..700 in one calcdist: the packed kernel picks 16 / 20 / 24 / 26 / 28 rows per lane per
    query pair and runs queries beyond 448 residues in several passes; every distance equals the oracle's."""
    rng = np.random.default_rng(11)
    lens = [int(x) for x in rng.integers(20, 700, size=56)] + [255, 256, 257, 320, 321, 384, 385, 416, 417, 448, 449, 450]
    enc = [rng.integers(3, 23, size=n).astype(np.uint8) for n in lens]
    M = np.array(golden("score_p24_blosum62")["matrix"])
    for sh in (-60, 12):
        prm = P.Params(P.ALPRM(sh=sh), vtype=1)
        want, _ = oracle.calcdist([oracle.seq(e) for e in enc], M, oracle.params(sh=sh, vtype=1))
        got = ctx.calcdist(P.SeqSet(enc), prm, M)
        assert np.array_equal(got, want), sh


def test_sharded_ranges_concatenate(ctx):
    seqs = gen_synth.config_set("c5a", 60)
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = np.array(golden("score_p24_blosum62")["matrix"])
    ss = P.SeqSet(enc)
    prm = P.Params()
    full = ctx.calcdist(ss, prm, M)
    npair = len(full)
    for world in (2, 3, 8):
        cuts = [npair * r // world for r in range(world + 1)]
        parts = [ctx.calcdist(ss, prm, M, cuts[r], cuts[r + 1]) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), full)
    assert len(ctx.calcdist(ss, prm, M, 7, 7)) == 0


def test_edge_cases_empty_and_tiny(ctx, oracle):
    rng = np.random.default_rng(3)
    lens = [0, 0, 1, 1, 2, 3, 15, 16, 17, 31, 32, 33, 511, 512, 513, 1025]
    enc = [rng.integers(3, 23, size=n).astype(np.uint8) for n in lens]
    M = np.array(golden("score_p24_blosum62")["matrix"])
    n = len(enc)
    ia = [i for j in range(1, n) for i in range(j)]
    ib = [j for j in range(1, n) for i in range(j)]
    for sh in (-60, 0, 5):
        prm = P.Params(P.ALPRM(sh=sh))
        op = oracle.params(sh=sh)
        got = ctx.score_pairs(P.SeqSet(enc), ia, ib, prm, M)
        want = np.array([oracle.aln_score_d(oracle.seq(enc[i]), oracle.seq(enc[j]), M, op) for i, j in zip(ia, ib)])
        assert np.array_equal(got.astype(np.float64), want), sh


def test_windows_left_right(ctx, oracle):
    rng = np.random.default_rng(4)
    enc = [rng.integers(3, 23, size=n).astype(np.uint8) for n in (120, 140, 90, 200)]
    left = np.array([0, 10, 5, 50], np.int32)
    right = np.array([120, 130, 90, 180], np.int32)
    M = np.array(golden("score_p24_blosum62")["matrix"])
    prm = P.Params()
    ss = P.SeqSet(enc, left=left, right=right)
    n = len(enc)
    ia = [i for j in range(1, n) for i in range(j)]
    ib = [j for j in range(1, n) for i in range(j)]
    got = ctx.score_pairs(ss, ia, ib, prm, M)
    op = oracle.params()
    want = np.array([oracle.aln_score_d(oracle.seq(enc[i], int(left[i]), int(right[i])),
                                        oracle.seq(enc[j], int(left[j]), int(right[j])), M, op)
                     for i, j in zip(ia, ib)])
    assert np.array_equal(got.astype(np.float64), want)


def test_full_size_properties_c2(ctx, oracle):
    """BASELINE config 2 at full size (1,000 x ~400 aa = 499,500 pairs): a seeded sample against the
    oracle, self-alignment = self score, and shard-invariance of a checksum."""
    seqs = gen_synth.config_set("c2")
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = np.array(golden("score_p24_blosum62")["matrix"])
    ss = P.SeqSet(enc)
    prm = P.Params()
    dist = ctx.calcdist(ss, prm, M)
    n = len(enc)
    assert len(dist) == n * (n - 1) // 2 and np.all(np.isfinite(dist))
    rng = np.random.default_rng(11)
    op = oracle.params()
    self_s = [oracle.lib().orc_self_score(__import__("ctypes").byref(oracle.seq(e)),
                                          oracle._mtx(M)[1], 25, __import__("ctypes").byref(op)) for e in enc]
    for _ in range(400):
        j = int(rng.integers(1, n))
        i = int(rng.integers(0, j))
        scr = oracle.aln_score_d(oracle.seq(enc[i]), oracle.seq(enc[j]), M, op)
        want = oracle.lib().orc_score2dist
        want.restype = __import__("ctypes").c_double
        C = __import__("ctypes")
        w = want(C.c_double(scr), C.c_int(len(enc[i])), C.c_int(len(enc[j])), C.c_double(self_s[i]),
                 C.c_double(self_s[j]), C.byref(op))
        assert float(dist[P.elem(i, j)]) == w, (i, j)
    # shard invariance: two halves computed separately hash to the same bytes
    half = len(dist) // 2
    a = ctx.calcdist(ss, prm, M, 0, half)
    b = ctx.calcdist(ss, prm, M, half, len(dist))
    assert np.array_equal(np.concatenate([a, b]), dist)
    # a sequence against itself scores its self score (no gap can beat the diagonal under BLOSUM62)
    idx = list(range(0, n, 50))
    sc = ctx.score_pairs(ss, idx, idx, prm, M)
    assert np.array_equal(sc.astype(np.float64), np.array([self_s[i] for i in idx]))


def test_fuzz_small_adversarial_pairs(ctx, oracle):
    """Thousands of tiny pairs (lengths 1..40, small alphabets, narrow bands, v down to 0) in one
    batch per parameter set: band cuts and boundary openings decide these scores."""
    M = np.array(golden("score_p24_blosum62")["matrix"])
    rng = np.random.default_rng(77)
    for sh, u, v in ((0, 1, 0), (1, 2, 9), (3, 3, 5), (-30, 2, 1), (-100, 1, 12), (2, 2, 9)):
        enc = []
        for _ in range(600):
            hi = 3 + int(rng.choice([2, 4, 20]))
            enc.append(rng.integers(3, hi, size=int(rng.integers(1, 40))).astype(np.uint8))
        ia = rng.integers(0, len(enc), size=3000).astype(np.int32)
        ib = rng.integers(0, len(enc), size=3000).astype(np.int32)
        prm = P.Params(P.ALPRM(u=u, v=v, sh=sh))
        got = ctx.score_pairs(P.SeqSet(enc), ia, ib, prm, M)
        op = oracle.params(u=u, v=v, sh=sh)
        want = np.array([oracle.aln_score_d(oracle.seq(enc[i]), oracle.seq(enc[j]), M, op) for i, j in zip(ia, ib)])
        assert np.array_equal(got.astype(np.float64), want), (sh, u, v)
