#!/usr/bin/env python
"""Config 5b (30 kb x 30 kb DNA pair): call time of pg_align_pairs, first call and repeats (debug aid)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import gen_synth  # noqa: E402
import prrn_aln_b200 as P  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402

dna = gen_synth.synth_set(2, 30000, 0.2, 0.2, 5, gen_synth.NT)
e2 = [seqcode.encode_dna(s) for s in dna]
Mn = np.full((18, 18), -4.0)
np.fill_diagonal(Mn, 2.0)
prm = P.Params(P.ALPRM(u=2, v=6, sh=-50))
ss = P.SeqSet(e2)
ctx = P.Context(0)
for k in range(4):
    t0 = time.perf_counter()
    sc, raw = ctx.align_pairs(ss, [0], [1], prm, Mn)
    dt = time.perf_counter() - t0
    print("call %d: %.2f ms, fill kernel %.2f ms, score %g, corners %d" % (k, dt * 1e3, ctx.last_kernel_ms(), sc[0], len(raw[0])))
ctx.close()
