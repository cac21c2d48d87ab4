#!/usr/bin/env python
"""K1F (floating-point alnScoreD) throughput on BASELINE config 2 with the reference's DEFAULT scoring
(PAM250, non-integral): calcdist(DynScr) over 1,000 x ~400 aa, float and double VTYPE, plus the
tgapf < 1 (lastD) and algmode.lcl (Fwd2d_vd + trimmed self scores) variants.  One JSON line each."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import gen_synth  # noqa: E402
import prrn_aln_b200 as P  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seqs = gen_synth.config_set(sys.argv[2] if len(sys.argv) > 2 else "c2", n)
    enc = [seqcode.encode_protein(s) for s in seqs]
    ss = P.SeqSet(enc)
    with open(os.path.join(ROOT, "tests", "golden", "score_p24_pam_f64.json")) as f:
        M = np.array(json.load(f)["matrix"])
    ctx = P.Context(0)
    for label, vt, tg, lcl in (("double", 1, 1.0, 0), ("float", 0, 1.0, 0), ("double tgapf=0.5", 1, 0.5, 0), ("double lcl=15", 1, 1.0, 15)):
        prm = P.Params(P.ALPRM(sh=-60, tgapf=tg), lcl=lcl, vtype=vt)
        cells = P.calcdist_cells(ss, prm)
        ctx.calcdist(ss, prm, M)
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            d = ctx.calcdist(ss, prm, M)
            t.append(time.perf_counter() - t0)
        print(json.dumps({"kernel": "k1f", "mode": label, "pairs": len(d), "cells": int(cells), "call_ms": 1e3 * min(t),
                          "gcups_e2e": cells / min(t) / 1e9, "checksum": float(np.sum(d))}))
    ctx.close()


if __name__ == "__main__":
    main()
