#!/usr/bin/env python
"""Secondary measurements (not the driver's bench contract): alignment-with-path throughput.

  all-pairs: `aln -ie` style alignment of the first N C2 sequences (pg_align_pairs: fill with
             direction bits + device traceback + D2H of the corner lists), GCUPS over band cells;
  long     : one DNA-like pair of L x L residues (multi-pass wavefront of a single warp).
Prints one JSON line per measurement.  CPU reference beside it when oracle/_ref exists.
"""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=200)
    ap.add_argument("--long", type=int, default=30000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    with open(os.path.join(ROOT, "tests", "golden", "score_p24_blosum62.json")) as f:
        M = np.array(json.load(f)["matrix"])
    ctx = P.Context(0)
    seqs = gen_synth.config_set("c2", a.n)
    enc = [seqcode.encode_protein(s) for s in seqs]
    ss = P.SeqSet(enc)
    n = len(enc)
    ia = [i for j in range(1, n) for i in range(j)]
    ib = [j for j in range(1, n) for i in range(j)]
    prm = P.Params(P.ALPRM(sh=-50))
    cells = P.calcdist_cells(ss, prm)
    ctx.align_pairs(ss, ia[:64], ib[:64], prm, M)
    best = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter()
        scores, raw = ctx.align_pairs(ss, ia, ib, prm, M)
        best = min(best, time.perf_counter() - t0)
    out = {"metric": "alignment-with-path GCUPS (all pairs, band cells, e2e incl. traceback + D2H)",
           "pairs": len(ia), "cells": cells, "seconds": best, "value": cells / best / 1e9, "unit": "GCUPS",
           "mean_corners": float(np.mean([len(r) for r in raw]))}
    if a.cpu:
        import refio
        if refio.available("f"):
            fa = "/tmp/bench_align.fa"
            m = min(a.n, 60)
            gen_synth.write_fasta(fa, seqs[:m])
            r = refio.run("align", fa, flavour="f", sh=-50)
            c2 = P.calcdist_cells(P.SeqSet(enc[:m]), prm)
            out["cpu_reference"] = {"value": c2 / r["time"] / 1e9, "unit": "GCUPS", "cores": 1,
                                    "sample": "first %d sequences, %d pairs, %.2f s" % (m, m * (m - 1) // 2, r["time"])}
    print(json.dumps(out))
    if a.long:
        dna = gen_synth.synth_set(2, a.long, 0.2, 0.2, 5, gen_synth.NT)
        e2 = [seqcode.encode_dna(s) for s in dna]
        Mn = np.full((17, 17), -4.0)
        np.fill_diagonal(Mn, 2.0)
        prm2 = P.Params(P.ALPRM(u=2, v=6, sh=-50))
        s2 = P.SeqSet(e2)
        cells2 = P.calcdist_cells(s2, prm2)
        t0 = time.perf_counter()
        sc, raw = ctx.align_pairs(s2, [0], [1], prm2, Mn)
        dt = time.perf_counter() - t0
        print(json.dumps({"metric": "single long pair alignment (one warp, multi-pass)", "len": [len(x) for x in dna],
                          "cells": cells2, "seconds": dt, "value": cells2 / dt / 1e9, "unit": "GCUPS",
                          "score": float(sc[0]), "corners": int(len(raw[0]))}))
    ctx.close()


if __name__ == "__main__":
    main()
