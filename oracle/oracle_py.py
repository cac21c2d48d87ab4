"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package (prrn_aln_b200 has no CPU fallback).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcSeq(C.Structure):
    _fields_ = [("res", C.POINTER(C.c_uint8)), ("len", C.c_int32), ("left", C.c_int32),
                ("right", C.c_int32), ("exgl", C.c_int32), ("exgr", C.c_int32)]


class OrcParams(C.Structure):
    _fields_ = [("u", C.c_double), ("v", C.c_double), ("scale", C.c_double), ("tgapf", C.c_double),
                ("u1", C.c_double), ("k1", C.c_int32), ("ls", C.c_int32), ("sh", C.c_int32),
                ("lcl", C.c_int32), ("vtype", C.c_int32)]


class OrcWindow(C.Structure):
    _fields_ = [("lw", C.c_int32), ("up", C.c_int32), ("width", C.c_int32)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_aln_score_d.restype = C.c_double
        L.orc_aln_score_d.argtypes = [C.POINTER(OrcSeq), C.POINTER(OrcSeq), C.POINTER(C.c_double),
                                      C.c_int, C.POINTER(OrcParams)]
        L.orc_self_score.restype = C.c_double
        L.orc_self_score.argtypes = [C.POINTER(OrcSeq), C.POINTER(C.c_double), C.c_int,
                                     C.POINTER(OrcParams)]
        L.orc_band_cells.restype = C.c_int64
        L.orc_band_cells.argtypes = [C.POINTER(OrcSeq), C.POINTER(OrcSeq), C.c_int]
        L.orc_stripe.restype = None
        L.orc_stripe.argtypes = [C.POINTER(OrcSeq), C.POINTER(OrcSeq), C.c_int, C.POINTER(OrcWindow)]
        L.orc_calcdist.restype = None
        L.orc_calcdist.argtypes = [C.POINTER(OrcSeq), C.c_int, C.POINTER(C.c_double), C.c_int,
                                   C.POINTER(OrcParams), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def params(u=2.0, v=9.0, scale=1.0, tgapf=1.0, u1=0.6, k1=7, ls=1, sh=-60, lcl=0, vtype=0):
    return OrcParams(u, v, scale, tgapf, u1, k1, ls, sh, lcl, vtype)


def seq(codes, left=None, right=None, exgl=0, exgr=0):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    s = OrcSeq(codes.ctypes.data_as(C.POINTER(C.c_uint8)), len(codes),
               0 if left is None else left, len(codes) if right is None else right, exgl, exgr)
    s._keep = codes
    return s


def _mtx(mtx):
    m = np.ascontiguousarray(mtx, dtype=np.float64)
    return m, m.ctypes.data_as(C.POINTER(C.c_double)), m.shape[0]


def aln_score_d(a, b, mtx, p):
    m, mp, dim = _mtx(mtx)
    return lib().orc_aln_score_d(C.byref(a), C.byref(b), mp, dim, C.byref(p))


def aln_score_full(a, b, mtx, p, want_ends=False):
    """alnScoreD with its dispatch (SWG when p.lcl & 16; Fwd2d_vd when want_ends): (score, ends|None)."""
    L = lib()
    L.orc_aln_score_full.restype = C.c_double
    L.orc_aln_score_full.argtypes = [C.POINTER(OrcSeq), C.POINTER(OrcSeq), C.POINTER(C.c_double), C.c_int,
                                     C.POINTER(OrcParams), C.POINTER(C.c_int)]
    m, mp, dim = _mtx(mtx)
    e = (C.c_int * 2)(0, 0)
    s = L.orc_aln_score_full(C.byref(a), C.byref(b), mp, dim, C.byref(p), e if want_ends else None)
    return s, ((e[0], e[1]) if want_ends else None)


def band_cells(a, b, sh):
    return lib().orc_band_cells(C.byref(a), C.byref(b), sh)


def stripe(a, b, sh):
    w = OrcWindow()
    lib().orc_stripe(C.byref(a), C.byref(b), sh, C.byref(w))
    return w.lw, w.up, w.width


def calcdist(seqs, mtx, p, want_scores=True):
    """seqs: list of OrcSeq.  Returns (dist[n(n-1)/2], raw scores) in elem(i,j) order."""
    n = len(seqs)
    arr = (OrcSeq * n)(*seqs)
    m, mp, dim = _mtx(mtx)
    npair = n * (n - 1) // 2
    dist = np.zeros(npair, dtype=np.float64)
    raw = np.zeros(npair, dtype=np.float64)
    lib().orc_calcdist(arr, n, mp, dim, C.byref(p), dist.ctypes.data_as(C.POINTER(C.c_double)),
                       raw.ctypes.data_as(C.POINTER(C.c_double)) if want_scores else None)
    return dist, raw


class OrcSkl(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32)]


def align_ngp(a, b, mtx, p, std=True):
    """alignC<DPunit> restatement (+ stdskl when std): returns (score, [(m, n), ...])."""
    L = lib()
    L.orc_align_ngp.restype = C.c_int
    L.orc_align_ngp.argtypes = [C.POINTER(OrcSeq), C.POINTER(OrcSeq), C.POINTER(C.c_double), C.c_int,
                                C.POINTER(OrcParams), C.POINTER(C.c_double), C.POINTER(OrcSkl), C.c_int]
    L.orc_stdskl.restype = C.c_int
    L.orc_stdskl.argtypes = [C.POINTER(OrcSkl), C.POINTER(OrcSkl)]
    m, mp, dim = _mtx(mtx)
    cap = 2 * (a.len + b.len) + 8
    out = (OrcSkl * cap)()
    scr = C.c_double(0)
    cnt = L.orc_align_ngp(C.byref(a), C.byref(b), mp, dim, C.byref(p), C.byref(scr), out, cap)
    if cnt < 0:
        raise RuntimeError("corner list overflow")
    if std:
        out2 = (OrcSkl * (2 * cnt + 4))()
        cnt = L.orc_stdskl(out, out2)
        out = out2
    return scr.value, [(out[i].m, out[i].n) for i in range(1, cnt + 1)]


# ---- group-to-group alignment (oracle_grp.c) ----------------------------------------------------
class OrcGfreq(C.Structure):
    _fields_ = [("glen", C.c_int32), ("freq", C.c_double), ("nres", C.c_int32)]


class OrcGroup(C.Structure):
    _fields_ = [("many", C.c_int32), ("len", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("nelm", C.c_int32), ("felm", C.c_int32), ("hetero", C.c_int32), ("nils", C.c_int32),
                ("cfq", C.c_void_p), ("dfq", C.c_void_p), ("efq", C.c_void_p), ("res", C.c_void_p),
                ("vss", C.c_void_p), ("weight", C.c_void_p), ("gpool", C.c_void_p),
                ("sfq", C.c_void_p), ("tfq", C.c_void_p), ("rfq", C.c_void_p)]


class OrcGparams(C.Structure):
    _fields_ = [("alnmode", C.c_int32), ("a_mode", C.c_int32), ("b_mode", C.c_int32), ("Noll", C.c_int32),
                ("codonk1", C.c_int32), ("sh", C.c_int32), ("vtype", C.c_int32), ("dxd", C.c_int32),
                ("u", C.c_double), ("Weighted_GOP", C.c_double), ("Basic_GOP", C.c_double),
                ("BasicGOP", C.c_double), ("BasicGEP", C.c_double), ("LongGOP", C.c_double), ("LongGEP", C.c_double)]


def group_arrays(g):
    """Flatten one parsed `galign` group (tools/refio.parse_galign) into the arrays orc_group / pg_group take."""
    npos = len(g["pos"])
    out = dict(many=g["many"], len=g["len"], left=g["left"], right=g["right"], nelm=g["nelm"], felm=g["felm"],
               hetero=g["hetero"], nils=g["nils"])
    out["cfq"] = np.array(g["cfq"], np.float64)
    out["dfq"] = np.array(g["dfq"], np.float64)
    out["efq"] = np.array(g["efq"], np.float64)
    out["res"] = np.ascontiguousarray(np.array(g["res"], np.uint8).reshape(npos, g["many"]))
    out["vss"] = np.ascontiguousarray(np.array(g["vss"], np.float64)) if g["nelm"] and len(g["vss"][0]) else None
    out["weight"] = np.array(g["weight"], np.float64) if g["weight"] else None
    pool = []
    offs = {}
    for tag in ("sfq", "tfq", "rfq"):
        o = np.full(npos, -1, np.int32)
        for x, lst in enumerate(g[tag]):
            if lst is None:
                continue
            o[x] = len(pool)
            pool.extend(lst)
            pool.append([-1, 0.0, 0])
        offs[tag] = o
    if not pool:
        pool = [[-1, 0.0, 0]]
    out["gpool"] = np.array([(int(a), float(b), int(c)) for a, b, c in pool],
                            dtype=np.dtype([("glen", np.int32), ("freq", np.float64), ("nres", np.int32)], align=True))
    out.update(offs)
    return out


def _orc_group(A):
    def p(x):
        return None if x is None else x.ctypes.data
    g = OrcGroup(A["many"], A["len"], A["left"], A["right"], A["nelm"], A["felm"], A["hetero"], A["nils"],
                 p(A["cfq"]), p(A["dfq"]), p(A["efq"]), p(A["res"]), p(A["vss"]), p(A["weight"]), p(A["gpool"]),
                 p(A["sfq"]), p(A["tfq"]), p(A["rfq"]))
    g._keep = A
    return g


def gparams_from_dump(d, sh=None):
    h, pm, pc = d["header"], d["pwdm"], d["pwdc"]
    return OrcGparams(pm["alnmode"], pm["a_mode"], pm["b_mode"], pm["Noll"], pm["codonk1"],
                      int(h["sh"]) if sh is None else sh, 1 if h["vtype"] == "f64" else 0, 1 if pm["DvsP"] == 0 else 0,
                      float(np.float32(float(h["u"]))), -float(np.float32(float(h["v"]))), pc["vgop1"],
                      pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])


def align_groups(A, B, mtx, gp):
    """alignC<recd_t> restatement on staged groups: (score, raw corner list, cells)."""
    L = lib()
    L.orc_align_groups.restype = C.c_int
    L.orc_align_groups.argtypes = [C.POINTER(OrcGroup), C.POINTER(OrcGroup), C.POINTER(C.c_double), C.c_int,
                                   C.POINTER(OrcGparams), C.POINTER(C.c_double), C.POINTER(OrcSkl), C.c_int,
                                   C.POINTER(C.c_int64)]
    m, mp, dim = _mtx(mtx)
    ga, gb = _orc_group(A), _orc_group(B)
    cap = 2 * (A["len"] + B["len"]) + 16
    out = (OrcSkl * cap)()
    scr = C.c_double(0)
    cells = C.c_int64(0)
    cnt = L.orc_align_groups(C.byref(ga), C.byref(gb), mp, dim, C.byref(gp), C.byref(scr), out, cap, C.byref(cells))
    if cnt < 0:
        raise RuntimeError("corner list overflow")
    return scr.value, [(out[i].m, out[i].n) for i in range(1, cnt + 1)], cells.value


def homscore_groups(A, B, mtx, gp):
    """HomScoreC<recd_t>(seqs, pwd, rr) restatement: (score, [rr0, rr1])."""
    L = lib()
    L.orc_homscore_groups.restype = C.c_int
    L.orc_homscore_groups.argtypes = [C.POINTER(OrcGroup), C.POINTER(OrcGroup), C.POINTER(C.c_double), C.c_int,
                                      C.POINTER(OrcGparams), C.POINTER(C.c_double), C.POINTER(C.c_long)]
    m, mp, dim = _mtx(mtx)
    ga, gb = _orc_group(A), _orc_group(B)
    scr = C.c_double(0)
    rr = (C.c_long * 2)()
    L.orc_homscore_groups(C.byref(ga), C.byref(gb), mp, dim, C.byref(gp), C.byref(scr), rr)
    return scr.value, [rr[0], rr[1]]


def swg_groups(A, B, mtx, gp):
    """swg1stC<SwgDPunit*> restatement (Fwd2c::forwardC, algmode.mlt <= 1): (best local score, dict of colony 0's
    box: mlb nlb mrb nrb lwr upr, cells)."""
    L = lib()
    L.orc_swg_groups.restype = C.c_int
    L.orc_swg_groups.argtypes = [C.POINTER(OrcGroup), C.POINTER(OrcGroup), C.POINTER(C.c_double), C.c_int,
                                 C.POINTER(OrcGparams), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    m, mp, dim = _mtx(mtx)
    ga, gb = _orc_group(A), _orc_group(B)
    val = C.c_double(0)
    box = (C.c_int * 6)()
    cells = C.c_int64(0)
    L.orc_swg_groups(C.byref(ga), C.byref(gb), mp, dim, C.byref(gp), C.byref(val), box, C.byref(cells))
    return val.value, dict(zip(("mlb", "nlb", "mrb", "nrb", "lwr", "upr"), [int(x) for x in box])), cells.value


def align_b1(a, b, mtx, p, std=True):
    """alignB_ng restatement (+ stdskl when std): (score, [(m, n), ...])."""
    L = lib()
    L.orc_align_b1.restype = C.c_int
    L.orc_align_b1.argtypes = [C.POINTER(OrcSeq), C.POINTER(OrcSeq), C.POINTER(C.c_double), C.c_int,
                               C.POINTER(OrcParams), C.POINTER(C.c_double), C.POINTER(OrcSkl), C.c_int]
    L.orc_stdskl.restype = C.c_int
    L.orc_stdskl.argtypes = [C.POINTER(OrcSkl), C.POINTER(OrcSkl)]
    m, mp, dim = _mtx(mtx)
    cap = 2 * (a.len + b.len) + 8
    out = (OrcSkl * cap)()
    scr = C.c_double(0)
    cnt = L.orc_align_b1(C.byref(a), C.byref(b), mp, dim, C.byref(p), C.byref(scr), out, cap)
    if cnt < 0:
        raise RuntimeError("orc_align_b1 failed (%d)" % cnt)
    if std:
        out2 = (OrcSkl * (2 * cnt + 4))()
        cnt = L.orc_stdskl(out, out2)
        out = out2
    return scr.value, [(out[i].m, out[i].n) for i in range(1, cnt + 1)]
