#!/usr/bin/env python
"""K1 (int32 DPX score kernel) with PG_FORCE_INT32=1 on C2- and C5a-like sets (debug aid for A/B runs: PRRN_GPU_LIB
picks the build, PG_K1_FIXED_ROWS=1 the fixed 16 rows per lane)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
os.environ["PG_FORCE_INT32"] = "1"
import numpy as np  # noqa: E402
import gen_synth  # noqa: E402
import prrn_aln_b200 as P  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402

M = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "score_p24_blosum62.json")))["matrix"])
prm = P.Params(P.ALPRM(sh=-60), vtype=1)
ctx = P.Context(0)
for cfg, n in (("c2", 1000), ("c5a", 2000)):
    enc = [seqcode.encode_protein(x) for x in gen_synth.config_set(cfg, n)]
    ss = P.SeqSet(enc)
    cells = P.calcdist_cells(ss, prm)
    ctx.calcdist(ss, prm, M)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        d = ctx.calcdist(ss, prm, M)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"kernel": "k1 int32", "set": cfg, "call_ms": best * 1e3, "gcups_e2e": cells / best / 1e9, "checksum": float(d.sum())}))
ctx.close()
