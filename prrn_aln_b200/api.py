"""ctypes binding of libprrn_gpu.so + the host-side mirror of the reference interface for the path.

Reference names are kept: alnScoreD (src/fwd2d1.cc:324), calcdist (src/phyl.cc:318), ALPRM
(src/seq.h:27), elem (src/cmn.h:115).  Errors follow the reference's convention of failing hard
(`fatal`): a non-zero return code from the C ABI raises PgError with the library's message.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PG_OK, PG_ERR_NO_DEVICE, PG_ERR_CUDA, PG_ERR_ARG, PG_ERR_UNSUPPORTED, PG_ERR_RANGE = range(6)


class PgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libprrn_gpu error %d: %s" % (code, msg))
        self.code = code


class ALPRM(C.Structure):
    """Mirror of struct ALPRM (reference src/seq.h:27-28); defaults of src/simmtx.cc:46 with the
    protein u, v of prrn5 (src/prrn5.cc:1262-1278)."""
    _fields_ = [("u", C.c_float), ("v", C.c_float), ("u0", C.c_float), ("u1", C.c_float),
                ("v0", C.c_float), ("tgapf", C.c_float), ("thr", C.c_float), ("scale", C.c_float),
                ("maxsp", C.c_float), ("gamma", C.c_float), ("k1", C.c_int32), ("ls", C.c_int32),
                ("sh", C.c_int32), ("mtx_no", C.c_int32)]

    def __init__(self, u=2.0, v=9.0, u0=0.0, u1=0.6, v0=0.0, tgapf=1.0, thr=35.0, scale=1.0,
                 maxsp=8.0, gamma=0.5, k1=7, ls=1, sh=-60, mtx_no=0):
        super().__init__(u, v, u0, u1, v0, tgapf, thr, scale, maxsp, gamma, k1, ls, sh, mtx_no)


class Params(C.Structure):
    _fields_ = [("alprm", ALPRM), ("lcl", C.c_int32), ("vtype", C.c_int32)]

    def __init__(self, alprm=None, lcl=0, vtype=0):
        super().__init__(alprm if alprm is not None else ALPRM(), lcl, vtype)

    @property
    def ftype(self):
        return np.float64 if self.vtype else np.float32


class _PgSeqs(C.Structure):
    _fields_ = [("res", C.c_void_p), ("offs", C.c_void_p), ("lens", C.c_void_p), ("left", C.c_void_p),
                ("right", C.c_void_p), ("exg", C.c_void_p), ("nseq", C.c_int32)]


class _PgGroup(C.Structure):
    _fields_ = [("many", C.c_int32), ("len", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("hetero", C.c_int32), ("nils", C.c_int32), ("cfq", C.c_void_p), ("efq", C.c_void_p),
                ("vec", C.c_void_p), ("glen", C.c_void_p), ("gfreq", C.c_void_p), ("npool", C.c_int32),
                ("sfq", C.c_void_p), ("tfq", C.c_void_p), ("rfq", C.c_void_p), ("gapmask", C.c_void_p),
                ("weight", C.c_void_p)]


class GParams(C.Structure):
    """pg_gparams: the PwdM / PwdB constants alignC reads (src/maln2.cc:227-243, src/aln2.cc:97-117)."""
    _fields_ = [("alnmode", C.c_int32), ("Noll", C.c_int32), ("codonk1", C.c_int32), ("sh", C.c_int32),
                ("kdim", C.c_int32), ("u", C.c_double), ("Weighted_GOP", C.c_double), ("Basic_GOP", C.c_double),
                ("BasicGOP", C.c_double), ("BasicGEP", C.c_double), ("LongGOP", C.c_double), ("LongGEP", C.c_double)]


def _pg_group(S):
    """S: dict of flat arrays from prrn_aln_b200.groups.stage_pair."""
    has = S["hetero"] >= 0
    g = _PgGroup(S["many"], S["len"], S["left"], S["right"], S["hetero"], S["nils"], S["cfq"].ctypes.data,
                 S["efq"].ctypes.data, S["vec"].ctypes.data, S["glen"].ctypes.data, S["gfreq"].ctypes.data,
                 len(S["glen"]), S["sfq"].ctypes.data if has else None, S["tfq"].ctypes.data if has else None,
                 S["rfq"].ctypes.data if has else None,
                 S["gapmask"].ctypes.data if S.get("gapmask") is not None else None,
                 S["weight"].ctypes.data if S.get("weight") is not None else None)
    return g


def lib_path():
    # PRRN_GPU_LIB: another build of the same library (A/B measurements); default: the in-tree build
    return os.environ.get("PRRN_GPU_LIB") or os.path.join(_HERE, "libprrn_gpu.so")


def declared_symbols():
    """Entry points declared in include/prrn_gpu.h (used by the CPU-side load test)."""
    hdr = os.path.join(os.path.dirname(_HERE), "include", "prrn_gpu.h")
    text = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def load_library():
    """Load libprrn_gpu.so; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise PgError(PG_ERR_NO_DEVICE, "libprrn_gpu.so is not built (run __graft_entry__.build()); "
                                         "there is no CPU fallback")
    L = C.CDLL(path)
    L.pg_version.restype = C.c_char_p
    L.pg_last_error.restype = C.c_char_p
    L.pg_last_error.argtypes = [C.c_void_p]
    L.pg_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.pg_destroy.argtypes = [C.c_void_p]
    L.pg_destroy.restype = None
    L.pg_score_pairs.argtypes = [C.c_void_p, C.POINTER(_PgSeqs), C.c_void_p, C.c_void_p, C.c_int64,
                                 C.POINTER(Params), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.pg_dist_pairs.argtypes = [C.c_void_p, C.POINTER(_PgSeqs), C.c_void_p, C.c_void_p, C.c_int64,
                                C.POINTER(Params), C.c_void_p, C.c_int32, C.c_void_p]
    L.pg_align_pairs.argtypes = [C.c_void_p, C.POINTER(_PgSeqs), C.c_void_p, C.c_void_p, C.c_int64,
                                 C.POINTER(Params), C.c_void_p, C.c_int32, C.c_void_p,
                                 C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int32))]
    L.pg_align_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int32))]
    L.pg_score_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.pg_local_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.pg_group_cells.restype = C.c_int64
    L.pg_group_cells.argtypes = [C.POINTER(_PgGroup), C.POINTER(_PgGroup), C.c_int32]
    L.pg_last_kernel_ms.restype = C.c_double
    L.pg_last_kernel_ms.argtypes = [C.c_void_p]
    L.pg_align_pairs_ng.argtypes = L.pg_align_pairs.argtypes
    L.pg_free.argtypes = [C.c_void_p]
    L.pg_free.restype = None
    L.pg_calcdist.argtypes = [C.c_void_p, C.POINTER(_PgSeqs), C.POINTER(Params), C.c_void_p, C.c_int32,
                              C.c_int64, C.c_int64, C.c_void_p]
    L.pg_seqs_upload.argtypes = [C.c_void_p, C.POINTER(_PgSeqs), C.POINTER(C.c_void_p)]
    L.pg_seqs_free.argtypes = [C.c_void_p, C.c_void_p]
    L.pg_seqs_free.restype = None
    L.pg_calcdist_dev.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int32,
                                  C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    L.pg_calcdist_cells.restype = C.c_int64
    L.pg_calcdist_cells.argtypes = [C.POINTER(_PgSeqs), C.POINTER(Params), C.c_int64, C.c_int64]
    L.pg_debug_packed_plan.argtypes = [C.POINTER(_PgSeqs), C.c_int64, C.c_int64, C.c_int32,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
    L.pg_dpx_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    _LIB = L
    return L


def elem(i, j):
    """Condensed index of the pair (i, j): reference src/cmn.h:115."""
    return j * (j - 1) // 2 + i if i < j else i * (i - 1) // 2 + j


class SeqSet:
    """A set of sequences in the layout the C ABI takes (concatenated residue codes + offsets)."""

    def __init__(self, encoded, left=None, right=None, exg=None):
        self.n = len(encoded)
        self.lens = np.array([len(e) for e in encoded], dtype=np.int32)
        self.offs = np.zeros(max(self.n, 1), dtype=np.int64)
        if self.n > 1:
            np.cumsum(self.lens[:-1], out=self.offs[1:self.n])
        self.res = (np.concatenate([np.asarray(e, dtype=np.uint8) for e in encoded])
                    if self.n and int(self.lens.sum()) else np.zeros(1, np.uint8))
        self.res = np.ascontiguousarray(self.res, dtype=np.uint8)
        self.left = None if left is None else np.ascontiguousarray(left, dtype=np.int32)
        self.right = None if right is None else np.ascontiguousarray(right, dtype=np.int32)
        self.exg = None if exg is None else np.ascontiguousarray(exg, dtype=np.uint8)

    def window_lens(self):
        l = self.left if self.left is not None else np.zeros(self.n, np.int32)
        r = self.right if self.right is not None else self.lens
        return (r - l).astype(np.int64)

    def c_struct(self):
        def p(a):
            return None if a is None else a.ctypes.data
        return _PgSeqs(p(self.res), p(self.offs), p(self.lens), p(self.left), p(self.right), p(self.exg), self.n)

    @property
    def nbytes(self):
        return int(self.res.nbytes + self.offs.nbytes + self.lens.nbytes)


def _mtx_for(prm, mtx):
    m = np.ascontiguousarray(mtx, dtype=prm.ftype)
    if m.ndim != 2 or m.shape[0] != m.shape[1]:
        raise ValueError("substitution matrix must be square")
    return m


class Context:
    """One CUDA device + workspace (pg_create / pg_destroy)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.pg_create(device, C.byref(h))
        if rc:
            raise PgError(rc, self.L.pg_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise PgError(rc, self.L.pg_last_error(self.h).decode())

    # -- per-call level: batch of alnScoreD -------------------------------------------------------
    def score_pairs(self, seqs, a_idx, b_idx, prm, mtx, want_ends=False):
        """Batch of alnScoreD(seqs[a], seqs[b]).  want_ends: also return the `ends` output (n x 2) of the
        semi-global variant (Fwd2d_vd, src/fwd2d1.cc:191-322)."""
        a = np.ascontiguousarray(a_idx, dtype=np.int32)
        b = np.ascontiguousarray(b_idx, dtype=np.int32)
        m = _mtx_for(prm, mtx)
        out = np.empty(len(a), dtype=prm.ftype)
        ends = np.zeros((len(a), 2), dtype=np.int32) if want_ends else None
        cs = seqs.c_struct()
        self._check(self.L.pg_score_pairs(self.h, C.byref(cs), a.ctypes.data, b.ctypes.data, len(a),
                                          C.byref(prm), m.ctypes.data, m.shape[0], out.ctypes.data,
                                          ends.ctypes.data if want_ends else None))
        return (out, ends) if want_ends else out

    def dist_pairs(self, seqs, a_idx, b_idx, prm, mtx):
        """100 * alnscore2dist(seqs[a], seqs[b]) for an explicit edge list (the DynScr branch of
        AdjacentMat::spaln_job, reference src/adjmat.cc:119-156): calcdist's distance for sparse candidate pairs."""
        a = np.ascontiguousarray(a_idx, dtype=np.int32)
        b = np.ascontiguousarray(b_idx, dtype=np.int32)
        m = _mtx_for(prm, mtx)
        out = np.empty(len(a), dtype=prm.ftype)
        cs = seqs.c_struct()
        self._check(self.L.pg_dist_pairs(self.h, C.byref(cs), a.ctypes.data, b.ctypes.data, len(a),
                                         C.byref(prm), m.ctypes.data, m.shape[0], out.ctypes.data))
        return out

    # -- per-call level: batch of alignC<DPunit> ------------------------------------------------
    def align_pairs(self, seqs, a_idx, b_idx, prm, mtx, ng=False):
        """Returns (scores, [corner array (n x 2, Vmf back-walk order) per pair]) as alignC does; ng=True:
        the Aln2b1 recurrence of alignB_ng / HomScoreB_ng (src/fwd2b1.cc) instead of Fwd2c<DPunit>."""
        a = np.ascontiguousarray(a_idx, dtype=np.int32)
        b = np.ascontiguousarray(b_idx, dtype=np.int32)
        m = _mtx_for(prm, mtx)
        out = np.empty(len(a), dtype=prm.ftype)
        offs = C.POINTER(C.c_int64)()
        pts = C.POINTER(C.c_int32)()
        cs = seqs.c_struct()
        fn = self.L.pg_align_pairs_ng if ng else self.L.pg_align_pairs
        self._check(fn(self.h, C.byref(cs), a.ctypes.data, b.ctypes.data, len(a), C.byref(prm),
                       m.ctypes.data, m.shape[0], out.ctypes.data, C.byref(offs), C.byref(pts)))
        try:
            o = np.ctypeslib.as_array(offs, shape=(len(a) + 1,)).copy()
            total = int(o[-1])
            p = np.ctypeslib.as_array(pts, shape=(max(total, 1) * 2,)).copy()[:2 * total].reshape(-1, 2)
        finally:
            self.L.pg_free(offs)
            self.L.pg_free(pts)
        return out, [p[o[i]:o[i + 1]] for i in range(len(a))]

    # -- per-call level, groups: batch of alignC<DPunit | DPunit_hf | DPunit_pf> --------------------
    def align_groups(self, pairs):
        """pairs: [(A, B, GParams), ...] with A / B from groups.stage_pair.  Returns (scores float64,
        [corner array (n x 2, Vmf back-walk order) per pair]) as alignC does for each pair."""
        n = len(pairs)
        ga = (_PgGroup * max(n, 1))(*[_pg_group(p[0]) for p in pairs])
        gb = (_PgGroup * max(n, 1))(*[_pg_group(p[1]) for p in pairs])
        gp = (GParams * max(n, 1))(*[p[2] for p in pairs])
        out = np.empty(n, np.float64)
        offs = C.POINTER(C.c_int64)()
        pts = C.POINTER(C.c_int32)()
        self._check(self.L.pg_align_groups(self.h, ga, gb, gp, n, out.ctypes.data, C.byref(offs), C.byref(pts)))
        try:
            o = np.ctypeslib.as_array(offs, shape=(n + 1,)).copy()
            total = int(o[-1])
            p = np.ctypeslib.as_array(pts, shape=(max(total, 1) * 2,)).copy()[:2 * total].reshape(-1, 2)
        finally:
            self.L.pg_free(offs)
            self.L.pg_free(pts)
        return out, [p[o[i]:o[i + 1]] for i in range(n)]

    def score_groups(self, pairs):
        """HomScoreC<recd_t>(seqs, pwd, rr) for a batch (reference src/fwd2c.h:663-668, dispatched by HomScore,
        src/maln2.cc:1837-1862): the same fill without the path store.  Returns (scores float64, rr int64 (n x 2))."""
        n = len(pairs)
        ga = (_PgGroup * max(n, 1))(*[_pg_group(p[0]) for p in pairs])
        gb = (_PgGroup * max(n, 1))(*[_pg_group(p[1]) for p in pairs])
        gp = (GParams * max(n, 1))(*[p[2] for p in pairs])
        out = np.empty(n, np.float64)
        rr = np.zeros((n, 2), np.int64)
        self._check(self.L.pg_score_groups(self.h, ga, gb, gp, n, out.ctypes.data, rr.ctypes.data))
        return out, rr

    def local_groups(self, pairs):
        """swg1stC<SwgDPunit*>(seqs, pwd) for a batch (reference src/fwd2c.h:697-701: Fwd2c::forwardC, algmode.mlt <= 1):
        Smith-Waterman on groups.  Returns (best local scores float64, boxes int32 (n x 6): mlb nlb mrb nrb lwr upr)."""
        n = len(pairs)
        ga = (_PgGroup * max(n, 1))(*[_pg_group(p[0]) for p in pairs])
        gb = (_PgGroup * max(n, 1))(*[_pg_group(p[1]) for p in pairs])
        gp = (GParams * max(n, 1))(*[p[2] for p in pairs])
        out = np.empty(n, np.float64)
        box = np.zeros((n, 6), np.int32)
        self._check(self.L.pg_local_groups(self.h, ga, gb, gp, n, out.ctypes.data, box.ctypes.data))
        return out, box

    # -- batch level: calcdist(DynScr) ------------------------------------------------------------
    def calcdist(self, seqs, prm, mtx, k_begin=0, k_end=None, out=None):
        npair = seqs.n * (seqs.n - 1) // 2
        k_end = npair if k_end is None else k_end
        m = _mtx_for(prm, mtx)
        if out is None:
            out = np.empty(max(k_end - k_begin, 0), dtype=prm.ftype)
        cs = seqs.c_struct()
        self._check(self.L.pg_calcdist(self.h, C.byref(cs), C.byref(prm), m.ctypes.data, m.shape[0],
                                       k_begin, k_end, out.ctypes.data))
        return out

    def upload(self, seqs):
        d = C.c_void_p()
        cs = seqs.c_struct()
        self._check(self.L.pg_seqs_upload(self.h, C.byref(cs), C.byref(d)))
        return d

    def free_seqs(self, d):
        self.L.pg_seqs_free(self.h, d)

    def calcdist_dev(self, dseqs, prm, mtx, k_begin, k_end, d_out_ptr, stream=None):
        """Asynchronous on `stream` (a raw cudaStream_t as int, or None); returns #kernel launches."""
        m = _mtx_for(prm, mtx)
        nl = C.c_int32(0)
        self._check(self.L.pg_calcdist_dev(self.h, dseqs, C.byref(prm), m.ctypes.data, m.shape[0], k_begin,
                                           k_end, C.c_void_p(d_out_ptr), C.c_void_p(stream or 0), C.byref(nl)))
        return nl.value

    def last_kernel_ms(self):
        return self.L.pg_last_kernel_ms(self.h)

    def dpx_peak(self):
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.L.pg_dpx_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def group_cells(A, B, sh):
    """DP cells the reference visits for one staged group pair (host-only helper)."""
    ga, gb = _pg_group(A), _pg_group(B)
    return load_library().pg_group_cells(C.byref(ga), C.byref(gb), sh)


def gparams_from_pwd(alnmode, Noll, codonk1, sh, kdim, u, v, Basic_GOP, BasicGOP, BasicGEP, LongGOP, LongGEP):
    """GParams from the PwdM / PwdB members a shim reads (Weighted_GOP = (VTYPE) -alnprm.v, maln2.cc:238)."""
    return GParams(alnmode, Noll, codonk1, sh, kdim, float(np.float32(u)), -float(np.float32(v)), Basic_GOP,
                   BasicGOP, BasicGEP, LongGOP, LongGEP)


def calcdist_cells(seqs, prm, k_begin=0, k_end=None):
    """DP cells the reference visits for the condensed range; host-only helper."""
    npair = seqs.n * (seqs.n - 1) // 2
    k_end = npair if k_end is None else k_end
    cs = seqs.c_struct()
    return load_library().pg_calcdist_cells(C.byref(cs), C.byref(prm), k_begin, k_end)


def packed_plan_coverage(seqs, k_begin=0, k_end=None, grid_blocks=444):
    """Host-only: (work items, (query pair, subject) slots, coverage count per condensed index) of
    the packed score kernel's schedule; every count must be 1."""
    npair = seqs.n * (seqs.n - 1) // 2
    k_end = npair if k_end is None else k_end
    cover = np.zeros(max(k_end - k_begin, 1), dtype=np.uint8)
    ni, ns = C.c_int64(0), C.c_int64(0)
    cs = seqs.c_struct()
    rc = load_library().pg_debug_packed_plan(C.byref(cs), k_begin, k_end, grid_blocks, C.byref(ni), C.byref(ns),
                                             cover.ctypes.data)
    if rc:
        raise PgError(rc, "pg_debug_packed_plan")
    return ni.value, ns.value, cover[:max(k_end - k_begin, 0)]


_DEFAULT_CTX = {}


def _ctx(device=0):
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]


def alnScoreD(seqs, sm, prm=None, pairs=None, device=0, ends=False):
    """VTYPE alnScoreD(const Seq* seqs[2], const Simmtx* sm, int* ends) -- reference
    src/fwd2d1.cc:324 -- for a batch.  `seqs` is a SeqSet; pairs = [(a, b), ...] (default: the two
    first sequences).  Returns the scores in the VTYPE of prm; with ends=True also the `ends` array
    (the reference's dispatch: prm.lcl & 16 -> Smith-Waterman-Gotoh score, ends -> Fwd2d_vd)."""
    prm = prm or Params()
    pairs = [(0, 1)] if pairs is None else pairs
    a = [p[0] for p in pairs]
    b = [p[1] for p in pairs]
    return _ctx(device).score_pairs(seqs, a, b, prm, sm, want_ends=ends)


def stdskl(corners):
    """SKL* stdskl(SKL** pskl) -- reference src/gaps.cc:139-175 -- the normalisation align2 applies to
    the corner list alignC returns: sort by (m, n), drop repeats / inconsistent points, interpolate
    the corner between a diagonal run and the gap that follows it."""
    pts = sorted((int(m), int(n)) for m, n in corners)
    if len(pts) < 2:
        return pts
    out = []
    pr = 2
    prv = pts[0]
    for cur in pts[1:]:
        dm, dn = cur[0] - prv[0], cur[1] - prv[1]
        if (not dm and not dn) or dm < 0 or dn < 0:
            continue
        dd = min(dm, dn)
        df = (dn > dm) - (dn < dm)
        if dd and df:
            if pr:
                out.append(prv)
            out.append((prv[0] + dd, prv[1] + dd))
        elif df != pr or not dm:
            out.append(prv)
        pr = df
        prv = cur
    out.append(prv)
    return out


def align2(seqs, sm, prm=None, pairs=None, device=0):
    """SKL* align2(mSeq* seqs[2], PwdM*, VTYPE* scr, Gsinfo*) -- reference src/maln2.cc:1875 -- for
    alnmode NGP_ALB (two single sequences): alignC<DPunit> on the GPU + stdskl.  Returns
    (scores, [normalised corner list per pair])."""
    prm = prm or Params()
    pairs = [(0, 1)] if pairs is None else pairs
    scores, raw = _ctx(device).align_pairs(seqs, [p[0] for p in pairs], [p[1] for p in pairs], prm, sm)
    return scores, [stdskl(r) for r in raw]


def alignB_ng(seqs, sm, prm=None, pairs=None, device=0):
    """SKL* alignB_ng(const Seq* seqs[2], const PwdB* pwd, VTYPE* scr) -- reference src/fwd2b1.cc:1347 --
    for a batch: Aln2b1 fill + traceback on the GPU, stdskl on the host.  Returns (scores, [corner list])."""
    prm = prm or Params()
    pairs = [(0, 1)] if pairs is None else pairs
    scores, raw = _ctx(device).align_pairs(seqs, [p[0] for p in pairs], [p[1] for p in pairs], prm, sm, ng=True)
    return scores, [stdskl(r) for r in raw]


def calcdist(seqs, sm, prm=None, device=0, k_begin=0, k_end=None):
    """FTYPE* calcdist(mSeq** sbuf, int nn, DynScr) -- reference src/phyl.cc:318 -- for single
    sequences: 100*(1 - (alnScoreD + u|dL|/2)/sqrt(self_i self_j)) in elem(i, j) order."""
    prm = prm or Params()
    return _ctx(device).calcdist(seqs, prm, sm, k_begin, k_end)
