#!/usr/bin/env python
"""Per-call latency of pg_align_groups on one small staged pair (fixed costs of the per-call binding)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import prrn_aln_b200 as P
from prrn_aln_b200 import groups as G

def load(name):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
    A, B = G.stage_pair(g["groups"][0], g["groups"][1], pm["a_mode"], pm["b_mode"], g["matrix"])
    gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                            float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
    return A, B, gp

ctx = P.Context(0)
for name in sys.argv[1:] or ["galign_gpf_prof12_raw5_wt"]:
    A, B, gp = load(name)
    for _ in range(5): ctx.align_groups([(A, B, gp)])
    n = 200
    t = time.perf_counter(); kms = 0.0
    for _ in range(n):
        ctx.align_groups([(A, B, gp)]); kms += ctx.last_kernel_ms()
    dt = (time.perf_counter() - t) / n
    print(json.dumps({"pair": name, "LQ": int(A["len"]), "LS": int(B["len"]), "call_us": dt * 1e6, "kernel_us": kms / n * 1e3}))
