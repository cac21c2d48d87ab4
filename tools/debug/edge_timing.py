#!/usr/bin/env python
"""Where the time of an edge-list distance call goes (debug aid): call time against edge count."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import json  # noqa: E402
import numpy as np  # noqa: E402
import gen_synth  # noqa: E402
import prrn_aln_b200 as P  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402

enc = [seqcode.encode_protein(x) for x in gen_synth.config_set("c5a")]
ss = P.SeqSet(enc)
n = len(enc)
M = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "score_p24_blosum62.json")))["matrix"])
prm = P.Params(P.ALPRM(sh=-60), vtype=1)
ctx = P.Context(0)
rng = np.random.default_rng(11)
for per in (1, 8, 32, 128):
    qi = np.repeat(np.arange(n), per).astype(np.int32)
    si = ((qi + 1 + rng.integers(0, n - 1, size=len(qi))) % n).astype(np.int32)
    ctx.dist_pairs(ss, qi, si, prm, M)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.dist_pairs(ss, qi, si, prm, M)
    dt = (time.perf_counter() - t0) / 3
    print("per query %4d edges %8d call %.2f ms" % (per, len(qi), dt * 1e3))
ctx.close()
