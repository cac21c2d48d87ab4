import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode
for name in ("align_p16_pam_f32", "align_p16_twopiece_default", "align_p16_twopiece_u1_1", "align_p16_twopiece_f64"):
    g = json.load(open(os.path.join(ROOT, "tests/golden/%s.json" % name)))
    h = g["params"]
    prm = P.Params(P.ALPRM(u=float(h["u"]), v=float(h["v"]), tgapf=float(h["tgapf"]), scale=float(h["scale"]), u1=float(h["u1"]), k1=int(h["k1"]), ls=int(h["ls"]), sh=int(h["sh"])), lcl=int(h["lcl"]), vtype=1 if h["vtype"] == "f64" else 0)
    enc = [seqcode.encode_protein(s) for s in g["seqs"]]
    ia = [p["i"] for p in g["pairs"]]; ib = [p["j"] for p in g["pairs"]]
    scores, raw = P.Context(0).align_pairs(P.SeqSet(enc), ia, ib, prm, np.array(g["matrix"]))
    nb = 0
    for k, p in enumerate(g["pairs"]):
        okp = P.stdskl(raw[k]) == [tuple(x) for x in p["skl"]]
        rel = abs(float(scores[k]) - p["score"]) / max(1, abs(p["score"]))
        if not okp or rel > 1e-6:
            nb += 1
            if nb < 6: print(name, p["i"], p["j"], "score", float(scores[k]), p["score"], "rel %.2e" % rel, "path same", okp)
    print(name, "pairs", len(g["pairs"]), "deviating", nb)
