"""Parity of kernel K3 (group-to-group alignC with gap-profile state, k3_groups.cu) through the C ABI
(pg_align_groups) against goldens frozen from the unmodified reference: the reference's own staged
inputs (ref_driver galign dump) go in, alignC's DP score and raw corner list must come out.
Tolerance (BASELINE north_star): scores within 1e-5 relative -- the kernel computes in double and
contracts sim2 in one order for all variants --, corner lists identical."""
import os
import sys

import numpy as np
import pytest

from conftest import golden, golden_names
import prrn_aln_b200 as P
from prrn_aln_b200 import groups as G

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

REL_TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = P.Context(0)
    yield c
    c.close()


def stage_golden(g, sh=None):
    pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
    A, B = G.stage_pair(g["groups"][0], g["groups"][1], pm["a_mode"], pm["b_mode"], g["matrix"], dxd=(pm["DvsP"] == 0))
    gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]) if sh is None else sh, A["vec"].shape[1],
                            float(h["u"]), float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
    return A, B, gp


@pytest.mark.parametrize("inline_sim", ["0", "1", "rl"])
@pytest.mark.parametrize("name", golden_names("galign_"))
def test_group_goldens(ctx, name, inline_sim, monkeypatch):
    """inline_sim 0: column score matrix precomputed by kernel K4 (FP64 tensor cores); 1: sim2 evaluated
    inside the DP cell (PG_K3_INLINE_SIM=1); rl: the register-list form of the cell (k3r_core.cuh, PG_K3_RL=1:
    fixed-capacity lists in registers, branch-free merges) for the gap-profile record modes.  All must reproduce
    the reference."""
    monkeypatch.setenv("PG_K3_INLINE_SIM", "1" if inline_sim == "1" else "0")
    monkeypatch.setenv("PG_K3_RL", "1" if inline_sim == "rl" else "0")
    g = golden(name)
    A, B, gp = stage_golden(g)
    scores, pts = ctx.align_groups([(A, B, gp)])
    want = g["alignc"]
    assert abs(scores[0] - want["score"]) <= REL_TOL * max(1.0, abs(want["score"]))
    assert pts[0].tolist() == want["skl"]
    lw, up, _ = g["window"]
    a, b = g["groups"]
    if g["pwdm"]["alnmode"] == G.NGP_ALN:       # the rectangle form (forwardA): no band
        return
    cells = sum(max(0, min(m + up + 1, b["right"]) - max(m + lw, b["left"])) for m in range(a["left"], a["right"]))
    assert P.group_cells(A, B, gp.sh) == cells


def test_homscore_goldens_in_one_call(ctx, oracle):
    """pg_score_groups = HomScoreC<recd_t>(seqs, pwd, rr) (src/fwd2c.h:663-668): the fill without the path store.
    Score and rr[2] against the reference's own HomScore for every golden pair (one batch), and against the oracle
    on other band shoulders."""
    gs = [golden(n) for n in golden_names("galign_") if "_rect_" not in n]
    scores, rr = ctx.score_groups([stage_golden(g) for g in gs])
    for k, g in enumerate(gs):
        want = g["homscore"]
        assert abs(scores[k] - want["score"]) <= REL_TOL * max(1.0, abs(want["score"])), g["name"]
        assert rr[k].tolist() == want["rr"], g["name"]
    for name in ("galign_ntv_3x1_wt", "galign_gpf_prof12_raw5_wt", "galign_gpf_highhetero"):
        g = golden(name)
        OA, OB = oracle.group_arrays(g["groups"][0]), oracle.group_arrays(g["groups"][1])
        for sh in (-100, -25, 3):
            A, B, gp = stage_golden(g, sh=sh)
            ws, wrr = oracle.homscore_groups(OA, OB, np.array(g["matrix"]), oracle.gparams_from_dump(g, sh=sh))
            s1, r1 = ctx.score_groups([(A, B, gp)])
            assert abs(s1[0] - ws) <= REL_TOL * max(1.0, abs(ws)) and r1[0].tolist() == wrr, (name, sh)


@pytest.mark.parametrize("tg", [None, "256", "128", "rl"])
def test_batch_of_all_goldens_in_one_call(ctx, tg, monkeypatch):
    """One launch over every golden pair (different modes, sizes, capacities), three times over: results must
    not depend on which CTA / arena a pair lands on -- with the default kernel (three threads per row) and with
    the one-thread-per-row variants (PG_K3_TG = 256 / 128 rows per alignment)."""
    monkeypatch.setenv("PG_K3_RL", "1" if tg == "rl" else "0")
    if tg is None or tg == "rl":
        monkeypatch.delenv("PG_K3_TG", raising=False)
    else:
        monkeypatch.setenv("PG_K3_TG", tg)
    gs = [golden(n) for n in golden_names("galign_")]
    staged = [stage_golden(g) for g in gs] * 3
    scores, pts = ctx.align_groups(staged)
    for k, (s, p) in enumerate(zip(scores, pts)):
        want = gs[k % len(gs)]["alignc"]
        assert abs(s - want["score"]) <= REL_TOL * max(1.0, abs(want["score"])), gs[k % len(gs)]["name"]
        assert p.tolist() == want["skl"], gs[k % len(gs)]["name"]


def test_matches_oracle_on_other_bands(ctx, oracle):
    """Same staged inputs, other band shoulders than the golden's: the oracle is the checker."""
    for name in ("galign_gpf_prof12_raw5_wt", "galign_hlf_prof10_single", "galign_gpf_twopiece", "galign_c1_multi_ab_f64"):
        g = golden(name)
        OA, OB = oracle.group_arrays(g["groups"][0]), oracle.group_arrays(g["groups"][1])
        for sh in (-100, -25, 3, 0):
            A, B, gp = stage_golden(g, sh=sh)
            want_s, want_p, cells = oracle.align_groups(OA, OB, np.array(g["matrix"]), oracle.gparams_from_dump(g, sh=sh))
            scores, pts = ctx.align_groups([(A, B, gp)])
            assert abs(scores[0] - want_s) <= REL_TOL * max(1.0, abs(want_s)), (name, sh)
            assert [tuple(x) for x in pts[0].tolist()] == want_p, (name, sh)
            assert P.group_cells(A, B, sh) == cells


def test_bad_arguments_fail_loudly(ctx):
    g = golden("galign_gpf_raw3x3")
    A, B, gp = stage_golden(g)
    gp.alnmode = 4          # GPF_ALN (rectangle with gap profiles): not built
    with pytest.raises(P.PgError) as e:
        ctx.align_groups([(A, B, gp)])
    assert e.value.code == 4
    # the rectangle form of HomScoreC (forwardA with island reports) is refused, not approximated by the banded one
    g = golden("galign_rect_single_p01")
    with pytest.raises(P.PgError) as e:
        ctx.score_groups([stage_golden(g)])
    assert e.value.code == 4


def test_edge_cases_groups(ctx):
    """Empty batch, tiny windows, a window inside the groups, refusal of over-long groups."""
    scores, pts = ctx.align_groups([])
    assert len(scores) == 0 and pts == []
    g = golden("galign_gpf_prof12_raw5_wt")
    A, B, gp = stage_golden(g)
    # one-column windows at the left end: the staged arrays keep their origin at left-1, so cut them
    def cut(S, right):
        T = dict(S)
        n = right - S["left"] + 1
        for k in ("cfq", "efq", "sfq", "tfq", "rfq"):
            T[k] = np.ascontiguousarray(S[k][:n])
        T["vec"] = np.ascontiguousarray(S["vec"][:n])
        T["right"] = right
        return T
    for ra, rb in ((A["left"] + 1, B["left"] + 1), (A["left"] + 1, B["left"] + 7), (A["left"] + 9, B["left"] + 1),
                   (A["left"] + 40, B["left"] + 33)):
        s1, p1 = ctx.align_groups([(cut(A, ra), cut(B, rb), gp)])
        assert np.isfinite(s1[0])
        q = p1[0].tolist()
        assert q[0] == [ra, rb] and q[-1] == [A["left"], B["left"]]          # back-walk order: end corner first
        assert all(q[i][0] >= q[i + 1][0] and q[i][1] >= q[i + 1][1] for i in range(len(q) - 1))
    A2 = dict(A)
    A2["len"] = 70000
    with pytest.raises(P.PgError) as e:
        ctx.align_groups([(A2, B, gp)])
    assert e.value.code == 5


def test_tiny_windows_match_oracle(ctx, oracle):
    """1..3-column windows of both groups against the oracle (boundary chains, first row / column rules)."""
    g = golden("galign_gpf_highhetero")
    base = oracle.gparams_from_dump(g)
    for la in (1, 2, 3):
        for lb in (1, 2, 3):
            gg = {"groups": [dict(g["groups"][0]), dict(g["groups"][1])]}
            for side, ln in ((0, la), (1, lb)):
                d = gg["groups"][side]
                n = ln + 1
                for k in ("pos", "cfq", "dfq", "efq", "res", "vss", "sfq", "tfq", "rfq"):
                    d[k] = d[k][:n]
                d["right"] = d["left"] + ln
            OA, OB = oracle.group_arrays(gg["groups"][0]), oracle.group_arrays(gg["groups"][1])
            want_s, want_p, _ = oracle.align_groups(OA, OB, np.array(g["matrix"]), base)
            pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
            A, B = G.stage_pair(gg["groups"][0], gg["groups"][1], pm["a_mode"], pm["b_mode"], g["matrix"])
            gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                    float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
            s1, p1 = ctx.align_groups([(A, B, gp)])
            assert abs(s1[0] - want_s) <= REL_TOL * max(1.0, abs(want_s)), (la, lb)
            assert [tuple(x) for x in p1[0].tolist()] == want_p, (la, lb)


def test_rectangle_form_matches_oracle_beyond_one_stripe(ctx, oracle, monkeypatch):
    """NGP_ALN (forwardA + initA, src/fwd2c.h:111-135,231-356): the rect goldens cover ~100-column inputs; here their
    columns are repeated 3..9 times so that the rows span several 256-row stripes / a thread-block cluster, alone and
    in one batch with banded pairs, against the oracle (itself pinned on the rect goldens)."""
    def tile(d, k):
        t = dict(d)
        n = d["right"] - d["left"]
        for key in ("pos", "cfq", "dfq", "efq", "res", "vss", "sfq", "tfq", "rfq"):
            v = d[key]
            t[key] = v[:1] + v[1:n + 1] * k + v[n + 1:]
        t["right"] = d["left"] + n * k
        t["len"] = d["len"] + n * (k - 1)
        return t
    cases = []
    for name, ka, kb in (("galign_rect_single_rag62", 3, 4), ("galign_rect_single_p24_twopiece", 9, 5),
                         ("galign_rect_ngp_gapless4x3_twopiece", 7, 7), ("galign_rect_ngp_gapless3x4_wt_f32", 4, 8)):
        g = golden(name)
        ga, gb = tile(g["groups"][0], ka), tile(g["groups"][1], kb)
        want_s, want_p, cells = oracle.align_groups(oracle.group_arrays(ga), oracle.group_arrays(gb), np.array(g["matrix"]),
                                                    oracle.gparams_from_dump(g))
        assert cells == (ga["right"] - ga["left"]) * (gb["right"] - gb["left"])
        pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
        A, B = G.stage_pair(ga, gb, pm["a_mode"], pm["b_mode"], g["matrix"])
        gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
        cases.append(((A, B, gp), want_s, want_p))
    assert max(c[0][0]["right"] for c in cases) > 600
    banded = [golden(n) for n in ("galign_gpf_twopiece", "galign_ngp_gapless4x3")]
    for cap in (None, "1"):
        if cap is None:
            monkeypatch.delenv("PG_K3_CLUSTER", raising=False)
        else:
            monkeypatch.setenv("PG_K3_CLUSTER", cap)
        for batch in ([c[0] for c in cases] + [stage_golden(g) for g in banded], [cases[1][0]]):
            scores, pts = ctx.align_groups(batch)
            for k, c in enumerate(cases if len(batch) > 1 else [cases[1]]):
                assert abs(scores[k] - c[1]) <= REL_TOL * max(1.0, abs(c[1])), (cap, k)
                assert [tuple(x) for x in pts[k].tolist()] == c[2], (cap, k)
            for k, g in enumerate(banded if len(batch) > 1 else []):
                w = g["alignc"]
                assert abs(scores[len(cases) + k] - w["score"]) <= REL_TOL * max(1.0, abs(w["score"]))
                assert pts[len(cases) + k].tolist() == w["skl"]
    # b must be staged from position 0 (forwardA's iterator starts there)
    A, B, gp = cases[0][0]
    B2 = dict(B)
    B2["left"] = 1
    with pytest.raises(P.PgError) as e:
        ctx.align_groups([(A, B2, gp)])
    assert e.value.code == 3


def test_smith_waterman_goldens_and_oracle(ctx, oracle):
    """pg_local_groups = swg1stC<SwgDPunit*> (Fwd2c::forwardC, algmode.mlt <= 1): the best local score and the box of
    colony 0 against the reference's own swg1st (8 goldens, one batch, three times over), then the second pass
    (swg2nd: align2 inside the box = pg_align_groups on that window) against the reference's swg2nd, and the oracle
    on inputs of several stripes (columns repeated) and other band shoulders."""
    gs = [golden(n) for n in golden_names("galign_swg_")]
    staged = [stage_golden(g) for g in gs]
    vals, boxes = ctx.local_groups(staged * 3)
    keys = ("mlb", "nlb", "mrb", "nrb", "lwr", "upr")
    for k in range(3 * len(gs)):
        w = gs[k % len(gs)]["swg"]
        assert abs(vals[k] - w["val"]) <= REL_TOL * max(1.0, abs(w["val"])), gs[k % len(gs)]["name"]
        assert boxes[k].tolist() == [w[q] for q in keys], gs[k % len(gs)]["name"]
    # second pass inside the box (swg2ndC swaps left / right with the colony's box and calls align2)
    batch, want = [], []
    for g, (A, B, gp), box in zip(gs, staged, boxes):
        if not g["swg2nd"]["skl"] or g["swg2nd"]["skl"][0] != [int(box[0]), int(box[1])]:
            continue                # align2 retried with another band: not the plain window call
        A2, B2 = dict(A), dict(B)
        for S, lo, hi in ((A2, int(box[0]), int(box[2])), (B2, int(box[1]), int(box[3]))):
            off = lo - S["left"]
            npos = hi - lo + 1
            for key in ("cfq", "efq", "sfq", "tfq", "rfq", "gapmask"):
                if S.get(key) is not None:
                    S[key] = np.ascontiguousarray(S[key][off:off + npos])
            S["vec"] = np.ascontiguousarray(S["vec"][off:off + npos])
            S["left"], S["right"] = lo, hi
        batch.append((A2, B2, gp))
        want.append(g)
    assert len(batch) >= 5
    scores, pts = ctx.align_groups(batch)
    for k, g in enumerate(want):
        assert abs(scores[k] - g["swg2nd"]["score"]) <= REL_TOL * max(1.0, abs(g["swg2nd"]["score"])), g["name"]
        assert P.stdskl(pts[k].tolist()) == [tuple(x) for x in g["swg2nd"]["skl"]], g["name"]

    def tile(d, k):
        t = dict(d)
        n = d["right"] - d["left"]
        for key in ("pos", "cfq", "dfq", "efq", "res", "vss", "sfq", "tfq", "rfq"):
            v = d[key]
            t[key] = v[:1] + v[1:n + 1] * k
        t["right"] = d["left"] + n * k
        t["len"] = d["len"] + n * (k - 1)
        return t
    for name, ka, kb, sh in (("galign_swg_single_unrelated", 3, 4, -60), ("galign_swg_gpf_twopiece", 4, 3, -25),
                             ("galign_swg_hlf_prof10_single", 5, 5, -100), ("galign_swg_ngp_gapless4x3", 4, 4, 10)):
        g = golden(name)
        ga, gb = tile(g["groups"][0], ka), tile(g["groups"][1], kb)
        wv, wb, _ = oracle.swg_groups(oracle.group_arrays(ga), oracle.group_arrays(gb), np.array(g["matrix"]),
                                      oracle.gparams_from_dump(g, sh=sh))
        pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
        A, B = G.stage_pair(ga, gb, pm["a_mode"], pm["b_mode"], g["matrix"], dxd=(pm["DvsP"] == 0))
        gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], sh, A["vec"].shape[1], float(h["u"]),
                                float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
        v1, b1 = ctx.local_groups([(A, B, gp)])
        assert abs(v1[0] - wv) <= REL_TOL * max(1.0, abs(wv)), name
        assert b1[0].tolist() == [wb[q] for q in keys], name
    with pytest.raises(P.PgError) as e:         # the rectangle form has no Smith-Waterman variant
        A, B, gp = stage_golden(golden("galign_rect_single_p01"))
        ctx.local_groups([(A, B, gp)])
    assert e.value.code == 4


def test_smith_waterman_and_rectangle_on_odd_windows(ctx, oracle):
    """Windows of 1 .. 300 columns cut out of tiled goldens (first / last stripe rows, single rows and columns, bands
    that leave a corner empty): pg_local_groups and the rectangle form against the oracle, all in two batches."""
    rng = np.random.default_rng(17)
    keys = ("mlb", "nlb", "mrb", "nrb", "lwr", "upr")

    def window(d, k, lo, ln):
        """columns lo .. lo+ln of the group tiled k times (array index 0 = position lo - 1)"""
        n = d["right"] - d["left"]
        t = dict(d)
        for key in ("pos", "cfq", "dfq", "efq", "res", "vss", "sfq", "tfq", "rfq"):
            v = d[key]
            full = v[:1] + v[1:n + 1] * k + v[n + 1:]
            t[key] = full[lo:lo + ln + 2]           # one column more than the banded forms read (the rectangle's b)
        t["left"], t["right"], t["len"] = 0, ln, ln
        return t
    sizes = [(1, 1), (1, 9), (8, 1), (2, 300), (257, 3), (33, 257), (300, 41), (129, 130), (64, 64), (255, 256)]
    swg_batch, swg_want, rect_batch, rect_want = [], [], [], []
    for name in ("galign_swg_single_unrelated", "galign_swg_gpf_twopiece", "galign_swg_hlf_prof10_single"):
        g = golden(name)
        pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
        for la, lb in sizes:
            ga = window(g["groups"][0], 5, int(rng.integers(0, 60)), la)
            gb = window(g["groups"][1], 5, int(rng.integers(0, 60)), lb)
            sh = int(rng.choice([-100, -60, -20, 0, 5]))
            wv, wb, _ = oracle.swg_groups(oracle.group_arrays(ga), oracle.group_arrays(gb), np.array(g["matrix"]),
                                          oracle.gparams_from_dump(g, sh=sh))
            A, B = G.stage_pair(ga, gb, pm["a_mode"], pm["b_mode"], g["matrix"], dxd=(pm["DvsP"] == 0))
            gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], sh, A["vec"].shape[1], float(h["u"]),
                                    float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
            swg_batch.append((A, B, gp))
            swg_want.append((wv, [wb[q] for q in keys], name, la, lb, sh))
    vals, boxes = ctx.local_groups(swg_batch)
    for k, (wv, wb, name, la, lb, sh) in enumerate(swg_want):
        assert abs(vals[k] - wv) <= REL_TOL * max(1.0, abs(wv)), (name, la, lb, sh)
        assert boxes[k].tolist() == wb, (name, la, lb, sh)
    for name in ("galign_rect_single_rag62", "galign_rect_ngp_gapless4x3_twopiece"):
        g = golden(name)
        pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
        for la, lb in sizes:
            ga = window(g["groups"][0], 5, int(rng.integers(0, 60)), la)
            gb = window(g["groups"][1], 5, int(rng.integers(0, 60)), lb)
            ws, wp, _ = oracle.align_groups(oracle.group_arrays(ga), oracle.group_arrays(gb), np.array(g["matrix"]),
                                            oracle.gparams_from_dump(g))
            A, B = G.stage_pair(ga, gb, pm["a_mode"], pm["b_mode"], g["matrix"])
            gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                    float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
            rect_batch.append((A, B, gp))
            rect_want.append((ws, wp, name, la, lb))
    scores, pts = ctx.align_groups(rect_batch)
    for k, (ws, wp, name, la, lb) in enumerate(rect_want):
        assert abs(scores[k] - ws) <= REL_TOL * max(1.0, abs(ws)), (name, la, lb)
        assert [tuple(x) for x in pts[k].tolist()] == wp, (name, la, lb)


def test_cluster_latency_kernel_matches_reference(ctx, monkeypatch):
    """Groups of ~1,100 columns in a latency-sized batch: K3 runs them on thread-block clusters (2 / 4 / 8 CTAs per
    alignment, rows handed from CTA to CTA through distributed shared memory).  Every score and corner list must
    equal the reference's own alignC (oracle/_ref/ref_driver_d, test infrastructure), with the cluster variant,
    with clusters capped at 2 CTAs and with the single-CTA kernel (PG_K3_CLUSTER=1)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import argparse
    import bench_groups
    import refio
    if not refio.available("d"):
        pytest.skip("oracle/_ref is not built")
    args = argparse.Namespace(members=30, length=1050, pairs=4, seed=5, sh=-60, cpu_rep=1)
    dumps = bench_groups.build_pairs(args)
    staged = []
    for d in dumps:
        pm, pc, h = d["pwdm"], d["pwdc"], d["header"]
        A, B = G.stage_pair(d["groups"][0], d["groups"][1], pm["a_mode"], pm["b_mode"], d["matrix"], dxd=(pm["DvsP"] == 0))
        gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
        staged.append((A, B, gp))
    assert max(int(s[0]["right"] - s[0]["left"]) for s in staged) > 256      # more rows than one CTA holds
    for cap in (None, "fenced", "release", "8", "2", "1"):
        monkeypatch.delenv("PG_K3_CLUSTER_FENCE", raising=False)
        if cap is None:                  # default: the records complete the barrier themselves (st.async + complete_tx)
            monkeypatch.delenv("PG_K3_CLUSTER", raising=False)
        elif cap == "fenced":            # release arrives + cluster-scope acquire in the per-step hand-shake
            monkeypatch.delenv("PG_K3_CLUSTER", raising=False)
            monkeypatch.setenv("PG_K3_CLUSTER_FENCE", "1")
        elif cap == "release":           # release arrives, CTA-scope waits (the default before the st.async form)
            monkeypatch.delenv("PG_K3_CLUSTER", raising=False)
            monkeypatch.setenv("PG_K3_CLUSTER_FENCE", "2")
        else:
            monkeypatch.setenv("PG_K3_CLUSTER", cap)
        for batch in (staged, staged[:1]):
            scores, pts = ctx.align_groups(batch)
            for k in range(len(batch)):
                w = dumps[k]["alignc"]
                assert abs(scores[k] - w["score"]) <= 1e-5 * max(1.0, abs(w["score"])), (cap, k)
                assert pts[k].tolist() == w["skl"], (cap, k)
