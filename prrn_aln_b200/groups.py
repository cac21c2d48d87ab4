"""Host-side staging of two groups for the group-to-group DP (pg_align_groups).

The reference stages a group for alignC inside PwdM::PwdM -> selAlnMode (src/maln2.cc:81-154,254-491):
mkthick (SeqThk per column), Gfq (gap-profile lists sfrq / tfrq / rfrq), convseq (frequency + profile
vector `vss`).  A shim that binds the C ABI walks those structures with mSeqItr, exactly like
oracle/ref_driver.cc's `galign` dump does; this module turns that per-column view into the flat arrays
`pg_group` takes, and folds every sim2 variant (sim11 .. sim33, src/maln2.cc:534-623,1230-1296) into
ONE contraction  S(m, n) = X_a[m] . Y_b[n]  over residue codes:

    a profile            : X_a = profile part of vss (pat),  Y_b = b's frequency vector
                           (profile b: frequency part of vss; raw b: sum of member weights per code)
    a raw, b profile     : X_a = a's frequency vector,       Y_b = b's profile part
    both raw             : X_a = sum_i w_i mtx[res_i][.],    Y_b = b's frequency vector
"""
import numpy as np

DECOMPACT = (0, 1, 2, 3, 5, 9)      # nil, gap, A, C, G, T (src/mseq.h:38) for sim33_n

# ALN_MODE (src/aln.h:71-76)
NGP_ALN = 1                         # the rectangle form (forwardA) of NGP_ALB; b's arrays hold one more column
NGP_ALB, HLF_ALB, RHF_ALB, GPF_ALB, NTV_ALB = 6, 7, 8, 9, 10
K3_MODE = {NGP_ALN: 0, NGP_ALB: 0, HLF_ALB: 1, RHF_ALB: 1, GPF_ALB: 2, NTV_ALB: 4}


def _onehot(g, dim):
    res = np.asarray(g["res"], np.int64)                 # [npos][many]
    w = np.ones(res.shape[1]) if g.get("weight") is None else np.asarray(g["weight"], np.float64)
    f = np.zeros((res.shape[0], dim))
    for i in range(res.shape[1]):
        np.add.at(f, (np.arange(res.shape[0]), res[:, i]), w[i])
    return f


def _pat(g, dim):
    v = np.asarray(g["vss"], np.float64)
    return np.ascontiguousarray(v[:, g["felm"]:g["felm"] + dim])


def _freqvec(g, dim, dxd):
    v = np.asarray(g["vss"], np.float64)
    f = np.zeros((v.shape[0], dim))
    if dxd:
        for j in range(g["felm"]):
            f[:, DECOMPACT[j]] = v[:, j]
    else:
        k = min(g["felm"], dim)
        f[:, :k] = v[:, :k]
    return f


def _lists(g):
    npos = len(g["cfq"])
    glen, freq = [], []
    offs = {}
    for tag in ("sfq", "tfq", "rfq"):
        o = np.full(npos, -1, np.int32)
        for x, lst in enumerate(g[tag]):
            if lst is None:
                continue
            o[x] = len(glen)
            for e in lst:
                glen.append(int(e[0]))
                freq.append(float(e[1]))
            glen.append(-1)
            freq.append(0.0)
        offs[tag] = o
    if not glen:
        glen, freq = [-1], [0.0]
    return np.array(glen, np.int32), np.array(freq, np.float64), offs


def stage_pair(ga, gb, a_mode, b_mode, mtx, dxd=False):
    """ga / gb: per-column views of the two groups (dicts with many, len, left, right, hetero, nils,
    cfq, efq, res, vss, nelm, felm, weight, sfq, tfq, rfq as in the reference's structures).
    Returns (A, B): dicts of flat arrays for pg_group."""
    mtx = np.nan_to_num(np.asarray(mtx, np.float64))
    dim = mtx.shape[0]
    if a_mode == 2:
        xa = _pat(ga, dim)
        yb = _freqvec(gb, dim, dxd) if b_mode == 2 else _onehot(gb, dim)
    elif b_mode == 2:
        xa = _onehot(ga, dim)
        yb = _pat(gb, dim)
    else:
        xa = _onehot(ga, dim) @ mtx
        yb = _onehot(gb, dim)
    out = []
    for g, vec in ((ga, xa), (gb, yb)):
        glen, gfreq, offs = _lists(g)
        res = np.asarray(g["res"], np.int64)
        gapmask = weight = None
        if res.ndim == 2 and res.shape[1] <= 32:     # NTV_ALB (DPunit_nv): IsGap bits per column + member weights
            gapmask = np.zeros(res.shape[0], np.uint32)
            for i in range(res.shape[1]):
                gapmask |= (res[:, i] <= 1).astype(np.uint32) << np.uint32(i)
            weight = (np.ones(res.shape[1]) if g.get("weight") is None else np.asarray(g["weight"], np.float64)).copy()
        out.append(dict(many=g["many"], len=g["len"], left=g["left"], right=g["right"], hetero=g["hetero"],
                        nils=g["nils"], cfq=np.ascontiguousarray(g["cfq"], np.float64),
                        efq=np.ascontiguousarray(g["efq"], np.float64), vec=np.ascontiguousarray(vec, np.float64),
                        glen=glen, gfreq=gfreq, sfq=offs["sfq"], tfq=offs["tfq"], rfq=offs["rfq"],
                        gapmask=gapmask, weight=weight))
    return out[0], out[1]
