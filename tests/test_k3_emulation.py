"""k3_core.cuh (the per-cell code of kernel K3: gap-profile lists, gapopen / update for DPunit,
DPunit_hf, DPunit_pf, two-piece states) driven by a host emulation of the CTA's anti-diagonal
wavefront (tests/host_emul/k3_emul.cc: slot rotation, band guards, pass hand-over, boundary chains,
path record store) against the goldens frozen from the reference.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, golden_names
from prrn_aln_b200 import groups as G


class K3Group(C.Structure):
    _fields_ = [("cfq", C.c_void_p), ("efq", C.c_void_p), ("prof", C.c_void_p), ("freq", C.c_void_p),
                ("glen", C.c_void_p), ("gfreq", C.c_void_p), ("sfq", C.c_void_p), ("tfq", C.c_void_p),
                ("rfq", C.c_void_p), ("L", C.c_int32), ("nils", C.c_int32), ("gapmask", C.c_void_p),
                ("weight", C.c_void_p), ("many", C.c_int32), ("pad", C.c_int32), ("blk", C.c_void_p)]


class K3Prm(C.Structure):
    _fields_ = [("mode", C.c_int32), ("Noll", C.c_int32), ("codonk1", C.c_int32), ("lw", C.c_int32), ("up", C.c_int32),
                ("capa", C.c_int32), ("capb", C.c_int32), ("kdim", C.c_int32), ("u", C.c_double), ("wgop", C.c_double),
                ("bgop", C.c_double), ("u2divu1", C.c_double), ("v2divv1", C.c_double), ("gop1", C.c_double),
                ("gep1", C.c_double), ("gop2", C.c_double), ("gep2", C.c_double), ("ltg_a", C.c_double), ("ltg_b", C.c_double),
                ("rtg_a", C.c_double), ("rtg_b", C.c_double), ("last_c", C.c_int32), ("last_r", C.c_int32),
                ("novmf", C.c_int32), ("origin_r", C.c_int32), ("rl", C.c_int32), ("rect", C.c_int32), ("swg", C.c_int32), ("pad3", C.c_int32)]


@pytest.fixture(scope="module")
def emul3():
    src = os.path.join(ROOT, "tests", "host_emul", "k3_emul.cc")
    out = os.path.join(ROOT, "tests", "host_emul", "libk3emul.so")
    subprocess.check_call(["g++", "-O2", "-fno-strict-aliasing", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    L = C.CDLL(out)
    L.k3_emul_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.k3_emul_align_rl.argtypes = L.k3_emul_align.argtypes + [C.c_int]
    return L


def _k3group(S):
    g = K3Group(S["cfq"].ctypes.data, S["efq"].ctypes.data, S["vec"].ctypes.data, S["vec"].ctypes.data,
                S["glen"].ctypes.data, S["gfreq"].ctypes.data, S["sfq"].ctypes.data, S["tfq"].ctypes.data,
                S["rfq"].ctypes.data, S["right"] - S["left"], S["nils"],
                S["gapmask"].ctypes.data if S.get("gapmask") is not None else None,
                S["weight"].ctypes.data if S.get("weight") is not None else None, S.get("many", 1), 0)
    g._keep = S
    return g


@pytest.mark.parametrize("name", golden_names("galign_"))
def test_wavefront_emulation_matches_reference(emul3, name):
    d = golden(name)
    pm, pc, h = d["pwdm"], d["pwdc"], d["header"]
    A, B = G.stage_pair(d["groups"][0], d["groups"][1], pm["a_mode"], pm["b_mode"], d["matrix"], dxd=(pm["DvsP"] == 0))
    lw, up, _ = d["window"]
    r0 = B["left"] - A["left"]
    mode = G.K3_MODE[pm["alnmode"]]
    capa, capb = max(A["hetero"], 0) + 3, max(B["hetero"], 0) + 3
    if mode == 4:
        capa, capb = (A["many"] + 1) // 2 + 1, (B["many"] + 1) // 2 + 1
    bgep, lgep, bgop, lgop = pc["BasicGEP"], pc["LongGEP"], pc["BasicGOP"], pc["LongGOP"]
    p = K3Prm(mode, pm["Noll"], pm["codonk1"], lw - r0, up - r0, capa, capb,
              A["vec"].shape[1], float(np.float32(float(h["u"]))), -float(np.float32(float(h["v"]))), pc["vgop1"],
              lgep / bgep if bgep < 0 else 0.0, lgop / bgop if bgop < 0 else 0.0, bgop, bgep, lgop, lgep, 1.0, 1.0)
    if pm["alnmode"] == G.NGP_ALN:      # forwardA: the whole rectangle; b staged one column further (position right)
        p.rect, p.lw, p.up = 1, -(A["right"] - A["left"]), B["right"] - B["left"]
        assert len(B["cfq"]) >= B["right"] - B["left"] + 2 and B["left"] == 0
    want = d["alignc"]
    for T in (256, 7, 33, -33, -5):     # rows per stripe: one stripe, many, ragged last; negative: reversed thread order
        out = np.zeros(2 * (A["len"] + B["len"] + 8), np.int32)
        sc = C.c_double(0)
        ga, gb = _k3group(A), _k3group(B)
        n = emul3.k3_emul_align(C.byref(ga), C.byref(gb), C.byref(p), T, A["left"], B["left"], C.byref(sc),
                                out.ctypes.data, len(out) // 2)
        assert n > 0, T
        assert abs(sc.value - want["score"]) <= 1e-5 * max(1.0, abs(want["score"])), T
        assert [[int(out[2 * i]), int(out[2 * i + 1])] for i in range(n)] == want["skl"], T
    if mode not in (1, 2):
        return
    # the register-list form of the cell (k3r_core.cuh): every capacity that holds this pair's lists must give the
    # very same score (bit for bit against the list-walking form) and corner list
    sc0 = C.c_double(0)
    out0 = np.zeros(2 * (A["len"] + B["len"] + 8), np.int32)
    ga, gb = _k3group(A), _k3group(B)
    emul3.k3_emul_align(C.byref(ga), C.byref(gb), C.byref(p), 33, A["left"], B["left"], C.byref(sc0), out0.ctypes.data, len(out0) // 2)
    need = max(A["hetero"], B["hetero"], 0)
    ran = 0
    for rl in (4, 6, 8):
        for T in (33, -7, 128):
            out = np.zeros(2 * (A["len"] + B["len"] + 8), np.int32)
            sc = C.c_double(0)
            n = emul3.k3_emul_align_rl(C.byref(ga), C.byref(gb), C.byref(p), T, A["left"], B["left"], C.byref(sc),
                                       out.ctypes.data, len(out) // 2, rl)
            if rl - 2 < need or n == -2:
                continue        # capacity below hetero or below the longest static list: the host never picks it
            assert n > 0, (rl, T, n)
            assert sc.value == sc0.value, (rl, T)
            assert [[int(out[2 * i]), int(out[2 * i + 1])] for i in range(n)] == want["skl"], (rl, T)
            ran += 1
    assert ran or need > 6


@pytest.mark.parametrize("name", golden_names("galign_swg_"))
def test_smith_waterman_wavefront_emulation(emul3, name):
    """The Smith-Waterman form of the K3 cell (k3s_* in k3_core.cuh: forwardC without secondary colonies) through the
    wavefront emulation: best local score and box of colony 0 against the reference's swg1st, for every stripe
    height and thread order (the per-thread maxima must reduce to the first maximum in row-major order)."""
    d = golden(name)
    pm, pc, h = d["pwdm"], d["pwdc"], d["header"]
    A, B = G.stage_pair(d["groups"][0], d["groups"][1], pm["a_mode"], pm["b_mode"], d["matrix"], dxd=(pm["DvsP"] == 0))
    lw, up, _ = d["window"]
    r0 = B["left"] - A["left"]
    mode = G.K3_MODE[pm["alnmode"]]
    capa, capb = max(A["hetero"], 0) + 3, max(B["hetero"], 0) + 3
    if mode == 4:
        capa, capb = (A["many"] + 1) // 2 + 1, (B["many"] + 1) // 2 + 1
    bgep, lgep, bgop, lgop = pc["BasicGEP"], pc["LongGEP"], pc["BasicGOP"], pc["LongGOP"]
    p = K3Prm(mode, pm["Noll"], pm["codonk1"], lw - r0, up - r0, capa, capb + 4,        # + 4: the box words
              A["vec"].shape[1], float(np.float32(float(h["u"]))), -float(np.float32(float(h["v"]))), pc["vgop1"],
              lgep / bgep if bgep < 0 else 0.0, lgop / bgop if bgop < 0 else 0.0, bgop, bgep, lgop, lgep, 1.0, 1.0)
    p.swg, p.origin_r = 1, r0
    want = d["swg"]
    for T in (256, 7, 33, -33, -5):
        out = np.zeros(16, np.int32)
        sc = C.c_double(0)
        ga, gb = _k3group(A), _k3group(B)
        n = emul3.k3_emul_align(C.byref(ga), C.byref(gb), C.byref(p), T, A["left"], B["left"], C.byref(sc), out.ctypes.data, 8)
        assert n == 3, T
        assert abs(sc.value - want["val"]) <= 1e-5 * max(1.0, abs(want["val"])), T
        assert [int(x) for x in out[:6]] == [want[k] for k in ("mlb", "nlb", "mrb", "nrb", "lwr", "upr")], T


def test_aln2b1_cell_emulation(emul3, oracle):
    """Mode 3 of the K3 machinery (Aln2b1::forwardB_ng cell, initB_ng chains) on groups of one."""
    from prrn_aln_b200 import seqcode
    import prrn_aln_b200 as P
    g = golden("alignb_p12_pam_f64")
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    M = np.nan_to_num(np.array(g["matrix"]))
    h = g["params"]
    fu, fv, fu1, k1 = float(h["u"]), float(h["v"]), float(h["u1"]), int(h["k1"])
    bgop, bgep, lgep = -fv, -fu, -float(np.float32(fu1))
    lgop = bgop - (lgep - bgep) * k1
    for pr in g["pairs"][:30]:
        a, b = enc[pr["i"]], enc[pr["j"]]
        la, lb = len(a), len(b)
        sides = []
        for e, asrow in ((a, True), (b, False)):
            npos = len(e) + 1
            vec = np.zeros((npos, 25))
            for x in range(1, npos):
                if asrow:
                    vec[x] = M[e[x - 1]]
                else:
                    vec[x, e[x - 1]] = 1.0
            sides.append(dict(cfq=np.ones(npos), efq=np.ones(npos), vec=np.ascontiguousarray(vec), glen=np.array([-1], np.int32),
                              gfreq=np.zeros(1), sfq=np.full(npos, -1, np.int32), tfq=np.full(npos, -1, np.int32),
                              rfq=np.full(npos, -1, np.int32), left=0, right=len(e), nils=0))
        lw, up, _ = oracle.stripe(oracle.seq(a), oracle.seq(b), int(h["sh"]))
        p = K3Prm(3, 2, 1 << 30, lw, up, 2, 2, 25, fu, -fv, -fv, 0.0, 0.0, bgop, bgep, lgop, lgep, 1.0, 1.0)
        out = np.zeros(2 * (la + lb + 8), np.int32)
        sc = C.c_double(0)
        ga, gb = _k3group(sides[0]), _k3group(sides[1])
        n = emul3.k3_emul_align(C.byref(ga), C.byref(gb), C.byref(p), 64, 0, 0, C.byref(sc), out.ctypes.data, len(out) // 2)
        assert sc.value == pr["score"], (pr["i"], pr["j"])
        assert P.stdskl([(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n)]) == [tuple(x) for x in pr["skl"]]
