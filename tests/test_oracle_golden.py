"""The oracle (oracle/oracle.c) against golden vectors frozen from the unmodified reference
(tools/make_golden.py -> tests/golden/*.json).  CPU only."""
import numpy as np
import pytest

from conftest import golden, golden_names
from prrn_aln_b200 import seqcode


def _oracle_params(O, g):
    h = g["params"]
    return O.params(u=float(h["u"]), v=float(h["v"]), scale=float(h["scale"]), tgapf=float(h["tgapf"]),
                    u1=float(h["u1"]), k1=int(h["k1"]), ls=int(h["ls"]), sh=int(h["sh"]),
                    lcl=int(h["lcl"]), vtype=1 if h["vtype"] == "f64" else 0)


def _exg_seq(O, e, gl, gr):
    """Seq::exg_seq (src/seq.cc:858-863) for a single ungapped sequence: only the inex flags change."""
    return O.seq(e, exgl=1 if gl else 0, exgr=1 if gr else 0)


@pytest.mark.parametrize("name", golden_names("score_"))
def test_scores_and_dist_bit_exact(oracle, name):
    g = golden(name)
    enc = [seqcode.encode_dna(s) if g["dna"] else seqcode.encode_protein(s) for s in g["seqs"]]
    p = _oracle_params(oracle, g)
    M = np.array(g["matrix"])
    lcl = p.lcl
    if lcl:
        # alnScoreD's other branches: swgforwardD (lcl & 16) and Fwd2d_vd with `ends` (the driver ran
        # exg_seq(lcl&1, lcl&2) on a and exg_seq(lcl&4, lcl&8) on b, as alnscore2dist does)
        k = 0
        for j in range(1, len(enc)):
            for i in range(j):
                if lcl & 16:
                    a, b = oracle.seq(enc[i]), oracle.seq(enc[j])
                else:
                    a, b = _exg_seq(oracle, enc[i], lcl & 1, lcl & 2), _exg_seq(oracle, enc[j], lcl & 4, lcl & 8)
                s, e = oracle.aln_score_full(a, b, M, p, want_ends="ends" in g)
                assert s == g["scores"][k], (i, j)
                if "ends" in g:
                    assert list(e) == g["ends"][k], (i, j)
                k += 1
        if "dist" in g:
            dist, _ = oracle.calcdist([oracle.seq(e) for e in enc], M, p)
            assert np.array_equal(dist, np.array(g["dist"])), "calcdist (lcl) vector differs from the reference"
        return
    dist, raw = oracle.calcdist([oracle.seq(e) for e in enc], M, p)
    assert np.array_equal(raw, np.array(g["scores"])), "alnScoreD scores differ from the reference"
    assert np.array_equal(dist, np.array(g["dist"])), "calcdist vector differs from the reference"


def test_band_cells_matches_loop_bounds(oracle):
    a = oracle.seq(np.zeros(10, np.uint8) + 3)
    b = oracle.seq(np.zeros(14, np.uint8) + 3)
    lw, up, width = oracle.stripe(a, b, -50)
    assert (lw, up, width) == (-5, 9, 17)
    cells = sum(min(m + up + 1, 14) - max(m + lw, 0) for m in range(10))
    assert oracle.band_cells(a, b, -50) == cells
    # absolute shoulder wider than the matrix -> full rectangle
    assert oracle.band_cells(a, b, 1000) == 140


@pytest.mark.parametrize("name", golden_names("align_"))
def test_alignment_scores_and_corner_lists_bit_exact(oracle, name):
    """orc_align_ngp + orc_stdskl against align2 (alignC<DPunit> + stdskl) of the reference."""
    g = golden(name)
    dna = g.get("args", {}).get("molc") == "n"      # long DNA pairs through aln's set-up (C5b, 6 kb)
    enc = [seqcode.encode_dna(s) if dna else seqcode.encode_protein(s) for s in g["seqs"]]
    p = _oracle_params(oracle, g)
    M = np.array(g["matrix"])
    for pr in g["pairs"]:
        scr, pts = oracle.align_ngp(oracle.seq(enc[pr["i"]]), oracle.seq(enc[pr["j"]]), M, p)
        assert scr == pr["score"], (pr["i"], pr["j"])
        assert pts == [tuple(x) for x in pr["skl"]], (pr["i"], pr["j"])


@pytest.mark.parametrize("name", golden_names("galign_"))
def test_group_alignment_bit_exact(oracle, name):
    """orc_align_groups (alignC<DPunit | DPunit_hf | DPunit_pf> restated) on the reference's own staged
    inputs: DP score and raw corner list identical to the reference, float and double flavours."""
    g = golden(name)
    A, B = oracle.group_arrays(g["groups"][0]), oracle.group_arrays(g["groups"][1])
    gp = oracle.gparams_from_dump(g)
    scr, pts, cells = oracle.align_groups(A, B, np.array(g["matrix"]), gp)
    assert scr == g["alignc"]["score"]
    assert pts == [tuple(x) for x in g["alignc"]["skl"]]
    lw, up, _ = g["window"]
    a, b = g["groups"]
    if g["pwdm"]["alnmode"] == 1:       # NGP_ALN: forwardA + initA over the whole rectangle (src/fwd2c.h:111-135,231-356)
        assert cells == (a["right"] - a["left"]) * (b["right"] - b["left"])
        return                          # (the rectangle form of HomScoreC, with its island reports, is not restated)
    want_cells = sum(max(0, min(m + up + 1, b["right"]) - max(m + lw, b["left"])) for m in range(a["left"], a["right"]))
    assert cells == want_cells
    # HomScore -> HomScoreC<recd_t>(seqs, pwd, rr): Fwd2c without Vmf, ptr = diagonal of the last first-row cell
    hs, rr = oracle.homscore_groups(A, B, np.array(g["matrix"]), gp)
    assert hs == g["homscore"]["score"] and rr == g["homscore"]["rr"]


@pytest.mark.parametrize("name", golden_names("galign_swg_"))
def test_group_smith_waterman_bit_exact(oracle, name):
    """orc_swg_groups (swg1stC<SwgDPunit*>: Fwd2c::initC + forwardC restated for algmode.mlt <= 1) on the reference's
    staged inputs: the best local score and the box of colony 0 identical to the reference's swg1st."""
    g = golden(name)
    A, B = oracle.group_arrays(g["groups"][0]), oracle.group_arrays(g["groups"][1])
    val, box, cells = oracle.swg_groups(A, B, np.array(g["matrix"]), oracle.gparams_from_dump(g))
    want = g["swg"]
    assert want["size"] == 0                    # mlt = 1: colony 0 only
    assert val == want["val"]
    assert box == {k: want[k] for k in box}
    lw, up, _ = g["window"]
    a, b = g["groups"]
    assert cells == sum(max(0, min(m + up + 1, b["right"]) - max(m + lw, b["left"])) for m in range(a["left"], a["right"]))


@pytest.mark.parametrize("name", golden_names("alignb_"))
def test_aln2b1_bit_exact(oracle, name):
    """orc_align_b1 (Aln2b1: alignB_ng + HomScoreB_ng restated) against the reference."""
    g = golden(name)
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    p = _oracle_params(oracle, g)
    M = np.nan_to_num(np.array(g["matrix"]))
    lcl = p.lcl                 # semi-global goldens: the driver ran exg_seq(lcl&1, lcl&2) on a, (lcl&4, lcl&8) on b
    for pr in g["pairs"]:
        scr, pts = oracle.align_b1(_exg_seq(oracle, enc[pr["i"]], lcl & 1, lcl & 2),
                                   _exg_seq(oracle, enc[pr["j"]], lcl & 4, lcl & 8), M, p)
        assert scr == pr["score"] == pr["hom"], (pr["i"], pr["j"])
        assert pts == [tuple(x) for x in pr["skl"]], (pr["i"], pr["j"])
