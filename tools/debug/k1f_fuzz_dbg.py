import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import json
import numpy as np
import prrn_aln_b200 as P
import oracle_py as oracle
M = np.nan_to_num(np.array(json.load(open(os.path.join(ROOT, "tests/golden/score_p24_pam_f32.json")))["matrix"]))
ctx = P.Context(0)
rng = np.random.default_rng(101)
for rep in range(6):
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
    nb = 0
    for k, (i, j) in enumerate(zip(ia, ib)):
        bad = float(sc[k]) != want[k][0] or (mode == 2 and tuple(r[1][k]) != want[k][1])
        if bad:
            nb += 1
            if nb <= 12:
                print("rep", rep, "mode", mode, "vt", vt, "sh", sh, "u", u, "v", v, "tg", tg, "| a len", len(enc[i]), left[i], right[i], exg[i],
                      "| b len", len(enc[j]), left[j], right[j], exg[j], "| got", float(sc[k]), r[1][k] if mode == 2 else "", "want", want[k])
    print("rep", rep, "mode", mode, "bad", nb, "of", len(ia))
