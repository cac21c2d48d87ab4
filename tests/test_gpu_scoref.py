"""Parity of kernel K1F (floating-point / semi-global / local alnScoreD, k1f_score.cu) with the
reference: goldens frozen from the unmodified reference (PAM matrices, tgapf < 1, algmode.lcl
variants incl. `ends` and the lcl branch of alnscore2dist) and seeded fuzz against the oracle.
Everything is compared BIT-EXACT: K1F performs the reference's IEEE operations in its order."""
import os
import sys

import numpy as np
import pytest

from conftest import golden, golden_names
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode

pytestmark = pytest.mark.gpu
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


def _pairs(n):
    return [i for j in range(1, n) for i in range(j)], [j for j in range(1, n) for i in range(j)]


@pytest.mark.parametrize("name", golden_names("score_"))
def test_all_score_goldens_through_float_kernel(ctx, name, monkeypatch):
    """Every score golden (integral ones too: PG_FORCE_FLOAT=1 keeps them off the DPX kernels)."""
    monkeypatch.setenv("PG_FORCE_FLOAT", "1")
    g = golden(name)
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    prm = _params(g)
    M = np.array(g["matrix"])
    n = len(enc)
    ia, ib = _pairs(n)
    lcl = prm.lcl
    if lcl and not lcl & 16:
        # the driver ran exg_seq(lcl&1, lcl&2) on a and exg_seq(lcl&4, lcl&8) on b per pair: build the
        # two roles as separate sequence sets (a-copies then b-copies)
        exg = np.array([lcl & 3] * n + [(lcl >> 2) & 3] * n, np.uint8)
        ss2 = P.SeqSet(enc + enc, exg=exg)
        sc, ends = ctx.score_pairs(ss2, ia, [j + n for j in ib], prm, M, want_ends=True)
        assert np.array_equal(sc.astype(np.float64), np.array(g["scores"]))
        assert np.array_equal(ends, np.array(g["ends"]))
    else:
        sc = ctx.score_pairs(P.SeqSet(enc), ia, ib, prm, M)
        assert np.array_equal(sc.astype(np.float64), np.array(g["scores"])), "alnScoreD differs from the reference"
    if "dist" in g:
        dist = ctx.calcdist(P.SeqSet(enc), prm, M)
        assert dist.dtype == prm.ftype
        assert np.array_equal(dist.astype(np.float64), np.array(g["dist"])), "calcdist differs from the reference"


@pytest.mark.parametrize("force_float", ["0", "1"])
@pytest.mark.parametrize("name", [n for n in golden_names("score_") if "dist" in golden(n)])
def test_edge_list_distances_equal_the_reference_calcdist(ctx, name, force_float, monkeypatch):
    """pg_dist_pairs (the DynScr branch of AdjacentMat::spaln_job, src/adjmat.cc:119-156: 100 * alnscore2dist per
    candidate pair) over ALL pairs of a golden set, in a shuffled order with repeats, must reproduce the reference's
    own calcdist vector bit for bit -- on the exact-integer kernels where they apply and on K1F, global and
    algmode.lcl branch (trimmed self scores)."""
    monkeypatch.setenv("PG_FORCE_FLOAT", force_float)
    g = golden(name)
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    prm = _params(g)
    M = np.array(g["matrix"])
    ia, ib = _pairs(len(enc))
    rng = np.random.default_rng(5)
    order = np.concatenate([rng.permutation(len(ia)), rng.integers(0, len(ia), size=17)])
    got = ctx.dist_pairs(P.SeqSet(enc), np.array(ia)[order], np.array(ib)[order], prm, M)
    assert got.dtype == prm.ftype
    assert np.array_equal(got.astype(np.float64), np.array(g["dist"])[order])
    assert len(ctx.dist_pairs(P.SeqSet(enc), [], [], prm, M)) == 0


def test_edge_list_distances_refuse_lcl16(ctx):
    g = golden("score_p24_lcl16")
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    with pytest.raises(P.PgError) as e:
        ctx.dist_pairs(P.SeqSet(enc), [0], [1], _params(g), np.array(g["matrix"]))
    assert e.value.code == 4


def test_calcdist_lcl16_is_refused(ctx):
    g = golden("score_p24_lcl16")
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    with pytest.raises(P.PgError) as e:
        ctx.calcdist(P.SeqSet(enc), _params(g), np.array(g["matrix"]))
    assert e.value.code == 4


def test_default_pam_routes_to_float_kernel_and_matches_oracle(ctx, oracle):
    """prrn's default scoring (PAM250, non-integral) on a C2 subset, float and double, with shards."""
    seqs = gen_synth.config_set("c2", 80)
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = np.array(golden("score_p24_pam_f64")["matrix"])
    for vt, sh in ((1, -60), (0, -60), (1, 15)):
        prm = P.Params(P.ALPRM(sh=sh), vtype=vt)
        want, _ = oracle.calcdist([oracle.seq(e) for e in enc], M, oracle.params(sh=sh, vtype=vt))
        got = ctx.calcdist(P.SeqSet(enc), prm, M)
        assert np.array_equal(got.astype(np.float64), want), (vt, sh)
        npair = len(want)
        parts = [ctx.calcdist(P.SeqSet(enc), prm, M, npair * r // 3, npair * (r + 1) // 3) for r in range(3)]
        assert np.array_equal(np.concatenate(parts).astype(np.float64), want)


def test_multipass_long_sequences(ctx, oracle):
    """Queries longer than one pass (512 rows float, 256 rows double / VD) in every mode."""
    g = golden("score_long1300")
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    enc += [enc[0][:300], enc[1][:513], enc[2][:257]]
    M = np.array(golden("score_p24_pam_f64")["matrix"])
    n = len(enc)
    ia, ib = _pairs(n)
    for vt in (0, 1):
        for lcl, tg, ends in ((0, 1.0, False), (0, 0.5, False), (16, 1.0, False), (0, 0.0, True)):
            prm = P.Params(P.ALPRM(tgapf=tg, sh=-40), lcl=lcl, vtype=vt)
            op = oracle.params(tgapf=tg, sh=-40, lcl=lcl, vtype=vt)
            r = ctx.score_pairs(P.SeqSet(enc), ia, ib, prm, M, want_ends=ends)
            want = [oracle.aln_score_full(oracle.seq(enc[i]), oracle.seq(enc[j]), M, op, want_ends=ends) for i, j in zip(ia, ib)]
            sc = r[0] if ends else r
            assert np.array_equal(sc.astype(np.float64), np.array([w[0] for w in want])), (vt, lcl, tg, ends)
            if ends:
                assert np.array_equal(r[1], np.array([w[1] for w in want])), (vt, lcl, tg)


def test_fuzz_modes_windows_flags(ctx, oracle):
    """Random small pairs with windows, end flags, tgapf, non-integral penalties, empty sequences, in
    all three modes, float and double: one batch per parameter set."""
    M = np.nan_to_num(np.array(golden("score_p24_pam_f32")["matrix"]))
    rng = np.random.default_rng(101)
    for rep in range(24):
        nseq = 300
        enc, left, right, exg = [], [], [], []
        for _ in range(nseq):
            hi = 3 + int(rng.choice([2, 4, 20]))
            ln = int(rng.integers(0, 45)) if rng.random() < 0.9 else int(rng.integers(100, 700))
            e = rng.integers(3, hi, size=ln).astype(np.uint8)
            l = int(rng.integers(0, ln + 1)) if rng.random() < 0.3 else 0
            r = int(rng.integers(l, ln + 1)) if rng.random() < 0.3 else ln
            enc.append(e); left.append(l); right.append(r)
            exg.append(int(rng.choice([0, 0, 0, 1, 2, 3])))
        mode = rep % 3
        vt = (rep // 3) % 2
        sh = int(rng.choice([-100, -60, -30, -10, 0, 1, 2, 3, 5, 100]))
        u, v = float(rng.choice([1, 2, 3, 0.6, 1.5])), float(rng.choice([0, 1, 5, 9, 12, 4.5]))
        tg = float(rng.choice([1, 0.5, 0, 0.3]))
        ss = P.SeqSet(enc, left=np.array(left, np.int32), right=np.array(right, np.int32), exg=np.array(exg, np.uint8))
        ia = rng.integers(0, nseq, size=1500).astype(np.int32)
        ib = rng.integers(0, nseq, size=1500).astype(np.int32)
        lcl = 16 if mode == 1 else 0
        prm = P.Params(P.ALPRM(u=u, v=v, sh=sh, tgapf=tg), lcl=lcl, vtype=vt)
        op = oracle.params(u=u, v=v, sh=sh, tgapf=tg, lcl=lcl, vtype=vt)
        r = ctx.score_pairs(ss, ia, ib, prm, M, want_ends=(mode == 2))
        sc = r[0] if mode == 2 else r

        def oseq(k):
            return oracle.seq(enc[k], left[k], right[k], exg[k] & 1, (exg[k] >> 1) & 1)
        want = [oracle.aln_score_full(oseq(i), oseq(j), M, op, want_ends=(mode == 2)) for i, j in zip(ia, ib)]
        assert np.array_equal(sc.astype(np.float64), np.array([w[0] for w in want])), (rep, mode, vt, sh, u, v, tg)
        if mode == 2:
            assert np.array_equal(r[1], np.array([w[1] for w in want])), (rep, vt, sh, u, v, tg)


def test_full_size_c2_default_scoring(ctx, oracle):
    """BASELINE config 2 at full size with the reference's DEFAULT scoring (PAM250, double VTYPE): 499,500
    pairs on K1F; a seeded sample against the oracle, shard invariance, finite values."""
    seqs = gen_synth.config_set("c2")
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = np.array(golden("score_p24_pam_f64")["matrix"])
    ss = P.SeqSet(enc)
    prm = P.Params(P.ALPRM(sh=-60), vtype=1)
    dist = ctx.calcdist(ss, prm, M)
    n = len(enc)
    assert len(dist) == n * (n - 1) // 2 and np.all(np.isfinite(dist))
    rng = np.random.default_rng(21)
    op = oracle.params(sh=-60, vtype=1)
    for _ in range(150):
        j = int(rng.integers(1, n))
        i = int(rng.integers(0, j))
        want, _ = oracle.calcdist([oracle.seq(enc[i]), oracle.seq(enc[j])], M, op)
        assert float(dist[P.elem(i, j)]) == want[0], (i, j)
    third = len(dist) // 3
    part = ctx.calcdist(ss, prm, M, third, third + 50000)
    assert np.array_equal(part, dist[third:third + 50000])
