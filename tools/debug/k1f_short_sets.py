#!/usr/bin/env python
"""K1F on sets of short sequences (debug aid: A/B of the rows-per-lane choice with PG_K1F_FIXED_ROWS=1)."""
import json
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

M = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "score_p24_pam_f64.json")))["matrix"])
ctx = P.Context(0)
for length, n in ((100, 4000), (180, 3000), (500, 1000)):
    enc = [seqcode.encode_protein(s) for s in gen_synth.synth_set(n, length, 0.1, 0.6, 3)]
    ss = P.SeqSet(enc)
    for vt in (0, 1):
        prm = P.Params(P.ALPRM(sh=-60), vtype=vt)
        cells = P.calcdist_cells(ss, prm)
        ctx.calcdist(ss, prm, M)
        t0 = time.perf_counter()
        d = ctx.calcdist(ss, prm, M)
        dt = time.perf_counter() - t0
        print(json.dumps({"len": length, "n": n, "vtype": vt, "gcups_e2e": cells / dt / 1e9, "checksum": float(np.sum(d, dtype=np.float64))}))
ctx.close()
