#!/usr/bin/env python
"""Group-to-group DP benchmark (BASELINE config 3 shape: partitions of a 200 x ~500 aa family).

Builds a synthetic family MSA, cuts it into candidate partitions (leaf edges: one member vs the rest ->
HLF/RHF; internal edges: two sub-alignments -> GPF), lets the reference stage each pair (PwdM) and
dumps what alignC reads (oracle/_ref/ref_driver_d galign: test infrastructure; it also times the
reference's own alignC on this box's host core), then runs all pairs in ONE pg_align_groups call.
Prints one JSON line: GCUPS (cells the reference visits), kernel-only and end-to-end, CPU beside it,
and checks every score / corner list against the reference.

  python tools/bench_groups.py --members 200 --length 500 --pairs 48 --cache /tmp/gb.pkl
"""
import argparse
import json
import os
import pickle
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import gen_msa  # noqa: E402
import refio  # noqa: E402


def build_pairs(args):
    fam = gen_msa.synth_msa(args.members, args.length, 0.1, 0.5, args.seed)
    rng = random.Random(args.seed + 1)
    n = len(fam)
    parts = []
    for k in range(args.pairs):
        if k % 3 == 0:                      # leaf edge
            i = rng.randrange(n)
            parts.append(([j for j in range(n) if j != i], [i]))
        else:                               # internal edge: contiguous block (the generator's order is a caterpillar tree)
            cut = rng.randrange(max(2, n // 8), n - max(2, n // 8))
            parts.append((list(range(cut)), list(range(cut, n))))
    out = []
    tmp = "/tmp/prrn_bench_groups_%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    for k, (ia, ib) in enumerate(parts):
        A, B = gen_msa.split_family(fam, ia, ib)
        fa, fb = os.path.join(tmp, "A%d" % k), os.path.join(tmp, "B%d" % k)
        gen_msa.write_native(fa, A, "A")
        gen_msa.write_native(fb, B, "B")
        d = refio.run_galign(fa, fb, flavour="d", wt=1, sh=args.sh, rep=args.cpu_rep)
        os.unlink(fa); os.unlink(fb)
        out.append(d)
        print("pair %d: alnmode %d, %d x %d members, %d x %d columns, hetero %d/%d, ref alignC %.1f ms" % (
            k, d["pwdm"]["alnmode"], d["groups"][0]["many"], d["groups"][1]["many"], d["groups"][0]["len"],
            d["groups"][1]["len"], d["groups"][0]["hetero"], d["groups"][1]["hetero"], 1e3 * d["time"]), file=sys.stderr)
    return out


def measure(dumps, replicate=16, steps=3, device=0, members=200, length=500, seed=7, sh=-60):
    """Run the staged pairs through pg_align_groups; returns the result dict (see module docstring)."""
    import prrn_aln_b200 as P
    from prrn_aln_b200 import groups as G
    staged = []
    cells = 0
    cpu_s = 0.0
    for d in dumps:
        pm, pc, h = d["pwdm"], d["pwdc"], d["header"]
        A, B = G.stage_pair(d["groups"][0], d["groups"][1], pm["a_mode"], pm["b_mode"], d["matrix"], dxd=(pm["DvsP"] == 0))
        gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
        staged.append((A, B, gp))
        cells += P.group_cells(A, B, gp.sh)
        cpu_s += d["time"]
    batch = staged * replicate
    ctx = P.Context(device)
    scores, pts = ctx.align_groups(batch)           # warm-up + parity
    bad = 0
    for k, d in enumerate(dumps):
        w = d["alignc"]
        if abs(scores[k] - w["score"]) > 1e-5 * max(1.0, abs(w["score"])) or pts[k].tolist() != w["skl"]:
            bad += 1
    for _ in range(2):
        ctx.align_groups(batch)
    kms, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        ctx.align_groups(batch)
        wall.append(time.perf_counter() - t0)
        kms.append(ctx.last_kernel_ms())
    ctx.close()
    tot_cells = cells * replicate
    return {"metric": "group-to-group DP GCUPS (alignC with gap-profile state, band cells)",
            "value": tot_cells / (np.median(kms) * 1e-3) / 1e9, "unit": "GCUPS",
            "e2e": {"value": tot_cells / np.median(wall) / 1e9, "unit": "GCUPS"},
            "kernel_ms": float(np.median(kms)), "call_ms": float(1e3 * np.median(wall)), "pairs": len(batch),
            "cells": int(tot_cells), "parity_mismatches": bad,
            "cpu_baseline": {"value": cells / cpu_s / 1e9, "unit": "GCUPS", "cores": 1, "kind": "reference",
                             "sample": "the same %d distinct pairs, alignC only, %.2f s" % (len(dumps), cpu_s)},
            "config": {"workload": "partitions of a synthetic %d x ~%d aa family (seed %d), sh=%d, PAM250 u=2 v=9, "
                                   "sequence weights on; %d distinct pairs x %d" % (members, length, seed, sh, len(dumps), replicate),
                       "modes": sorted(set(d["pwdm"]["alnmode"] for d in dumps)),
                       "mean_hetero": float(np.mean([max(g["hetero"], 0) for d in dumps for g in d["groups"]]))}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=200)
    ap.add_argument("--length", type=int, default=500)
    ap.add_argument("--pairs", type=int, default=48)
    ap.add_argument("--sh", type=int, default=-60)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-rep", type=int, default=1)
    ap.add_argument("--replicate", type=int, default=1, help="run each pair this many times in the batch")
    ap.add_argument("--cache", default=None, help="pickle of the staged dumps (built where oracle/_ref runs)")
    ap.add_argument("--stage-only", action="store_true")
    args = ap.parse_args()
    if args.cache and os.path.exists(args.cache):
        dumps = pickle.load(open(args.cache, "rb"))
    else:
        dumps = build_pairs(args)
        if args.cache:
            pickle.dump(dumps, open(args.cache, "wb"))
    if args.stage_only:
        return
    print(json.dumps(measure(dumps, args.replicate, args.steps, 0, args.members, args.length, args.seed, args.sh)))


if __name__ == "__main__":
    main()
