"""Fill time of the striped long-pair kernel (K2) over pair lengths; PG_K2_LONG_ROWS / PG_K2_CHUNK select variants."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import gen_synth
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode

ctx = P.Context(0)
Mn = np.full((18, 18), -4.0)
np.fill_diagonal(Mn, 2.0)
prm = P.Params(P.ALPRM(u=2, v=6, sh=-50))
for length in [int(x) for x in (sys.argv[1:] or ["3000", "6000", "12000", "20000", "30000", "45000"])]:
    dna = gen_synth.synth_set(2, length, 0.2, 0.2, 5, gen_synth.NT)
    ss = P.SeqSet([seqcode.encode_dna(s) for s in dna])
    cells = P.calcdist_cells(ss, prm)
    ctx.align_pairs(ss, [0], [1], prm, Mn)
    t = []
    f = []
    for _ in range(3):
        t0 = time.perf_counter()
        sc, raw = ctx.align_pairs(ss, [0], [1], prm, Mn)
        t.append(time.perf_counter() - t0)
        f.append(ctx.last_kernel_ms())
    print(json.dumps({"len": length, "cells": int(cells), "fill_ms": min(f), "call_ms": 1e3 * min(t), "gcups_fill": cells / min(f) / 1e6,
                      "score": float(sc[0]), "corners": len(raw[0]), "rows": os.environ.get("PG_K2_LONG_ROWS", "default"),
                      "chunk": os.environ.get("PG_K2_CHUNK", "default")}))
