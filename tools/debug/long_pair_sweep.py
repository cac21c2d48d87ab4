#!/usr/bin/env python
"""Long DNA pairs of several lengths through pg_align_pairs (debug aid): fill kernel time and call time; the path is
re-scored on the host.  PG_K2_LONG_ROWS / PG_K2_WIDE / PG_K2_LONG_V1 select the kernel form."""
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

Mn = np.full((18, 18), -4.0)
np.fill_diagonal(Mn, 2.0)
prm = P.Params(P.ALPRM(u=2, v=6, sh=-50))
ctx = P.Context(0)
for length in (6000, 12000, 20000, 30000, 45000):
    dna = gen_synth.synth_set(2, length, 0.2, 0.2, 5, gen_synth.NT)
    e2 = [seqcode.encode_dna(s) for s in dna]
    ss = P.SeqSet(e2)
    ctx.align_pairs(ss, [0], [1], prm, Mn)
    t0 = time.perf_counter()
    sc, raw = ctx.align_pairs(ss, [0], [1], prm, Mn)
    dt = time.perf_counter() - t0
    pts = P.stdskl(raw[0])
    a, b = e2
    s = 0.0
    for (m0, n0), (m1, n1) in zip(pts[:-1], pts[1:]):
        dm, dn = m1 - m0, n1 - n0
        s += float(np.sum(Mn[a[m0:m1], b[n0:n1]])) if dm == dn else -(6 + 2 * (dm + dn))
    print("len %6d: call %.2f ms, fill %.2f ms, score %g, rescored %g, corners %d" % (length, dt * 1e3, ctx.last_kernel_ms(), sc[0], s, len(pts)))
ctx.close()
