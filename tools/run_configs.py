#!/usr/bin/env python
"""BASELINE.json configs beyond the bench line, run once per round on the GPU box (one JSON line each):
  c5a  all-vs-all calcdist of 10,000 x ~300 aa (49,995,000 pairs): throughput, a seeded sample against
       the oracle, shard invariance;
  c5b  one DNA pair of 30 kb x 30 kb: alignC<DPunit> with path (K2), path re-scored on the host;
  c4   partitions of a DNA family of ~2 kb with two-piece gap penalties (-yl3): the reference stages and
       aligns each pair (oracle/_ref/ref_driver_d galign), K4 + K3 must return its score / corner list.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import gen_msa  # noqa: E402
import gen_synth  # noqa: E402
import prrn_aln_b200 as P  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402
from prrn_aln_b200 import groups as G  # noqa: E402


def blosum():
    with open(os.path.join(ROOT, "tests", "golden", "score_p24_blosum62.json")) as f:
        return np.array(json.load(f)["matrix"])


def c5a(ctx, n=10000):
    import oracle_py as O
    seqs = gen_synth.config_set("c5a", n)
    enc = [seqcode.encode_protein(s) for s in seqs]
    ss = P.SeqSet(enc)
    M = blosum()
    prm = P.Params(P.ALPRM(sh=-60), vtype=1)
    cells = P.calcdist_cells(ss, prm)
    ctx.calcdist(ss, prm, M, 0, 1000)
    t0 = time.perf_counter()
    d = ctx.calcdist(ss, prm, M)
    dt = time.perf_counter() - t0
    rng = np.random.default_rng(3)
    op = O.params(sh=-60, vtype=1)
    bad = 0
    for _ in range(300):
        j = int(rng.integers(1, len(enc)))
        i = int(rng.integers(0, j))
        want, _ = O.calcdist([O.seq(enc[i]), O.seq(enc[j])], M, op)
        bad += float(d[P.elem(i, j)]) != want[0]
    npair = len(d)
    h = npair // 3
    part = ctx.calcdist(ss, prm, M, h, h + 100000)
    bad += not np.array_equal(part, d[h:h + 100000])
    out = {"config": "c5a", "sequences": len(enc), "pairs": npair, "cells": int(cells), "seconds_e2e": dt,
           "gcups_e2e": cells / dt / 1e9, "oracle_sample_mismatches": int(bad), "checksum": float(d.sum())}
    # CPU beside it: the reference's own calcdist on a 1 % sample of the pairs (every 10th sequence: 1,000 sequences,
    # 499,500 pairs of the same length distribution) on all host cores, compared value by value, extrapolated by cells
    import refio
    if refio.available("d"):
        from prrn_aln_b200 import sharding
        sub = list(range(0, len(seqs), 10))
        fa = "/tmp/prrn_c5a_sample_%d.fa" % os.getpid()
        gen_synth.write_fasta(fa, [seqs[i] for i in sub])
        th = os.cpu_count() or 1
        r = refio.run("dist", fa, flavour="d", threads=th, sh=-60)
        os.unlink(fa)
        want = np.array(r["dist"])
        got = np.array([d[P.elem(sub[i], sub[j])] for j in range(1, len(sub)) for i in range(j)])
        scells = int(sharding.row_costs(np.array([len(seqs[i]) for i in sub]), -60)[0].sum())
        out["cpu_sample"] = {"kind": "reference", "cores": th, "pairs": len(want), "cells": scells, "seconds": r["time"],
                             "gcups": scells / r["time"] / 1e9, "mismatches_vs_gpu": int(np.sum(want != got)),
                             "extrapolated_seconds_all_pairs": r["time"] * cells / scells}
    print(json.dumps(out))


def c5b(ctx, length=30000):
    dna = gen_synth.synth_set(2, length, 0.2, 0.2, 5, gen_synth.NT)
    e2 = [seqcode.encode_dna(s) for s in dna]
    Mn = np.full((18, 18), -4.0)
    np.fill_diagonal(Mn, 2.0)
    prm = P.Params(P.ALPRM(u=2, v=6, sh=-50))
    ss = P.SeqSet(e2)
    cells = P.calcdist_cells(ss, prm)
    t0 = time.perf_counter()
    sc, raw = ctx.align_pairs(ss, [0], [1], prm, Mn)
    dt_first = time.perf_counter() - t0         # first call: grows the context's direction-bit store (420 MB)
    t0 = time.perf_counter()
    sc, raw = ctx.align_pairs(ss, [0], [1], prm, Mn)
    dt = time.perf_counter() - t0
    fill_ms = ctx.last_kernel_ms()
    pts = P.stdskl(raw[0])
    a, b = e2
    s = 0.0
    ok = pts[0] == (0, 0) and pts[-1] == (len(a), len(b))
    for (m0, n0), (m1, n1) in zip(pts[:-1], pts[1:]):
        dm, dn = m1 - m0, n1 - n0
        ok = ok and dm >= 0 and dn >= 0 and (dm == dn or dm == 0 or dn == 0)
        if dm == dn:
            s += float(np.sum(Mn[a[m0:m1], b[n0:n1]]))
        else:
            s -= 6 + 2 * (dm + dn)
    out = {"config": "c5b", "len": [len(a), len(b)], "cells": int(cells), "seconds_e2e": dt, "seconds_first_call": dt_first,
           "gcups_e2e": cells / dt / 1e9,
           "fill_kernel_ms": fill_ms, "gcups_fill": cells / (fill_ms * 1e-3) / 1e9, "score": float(sc[0]), "path_rescored": s,
           "path_valid": bool(ok and s == float(sc[0])), "corners": len(pts)}
    gpath = os.path.join(ROOT, "tests", "golden", "align_c5b_30k.json")
    if length == 30000 and os.path.exists(gpath):       # the reference's own corner list for this pair (tools/make_golden.py dna)
        g = json.load(open(gpath))["pairs"][0]
        out["equals_reference_alignment"] = bool(float(sc[0]) == g["score"] and pts == [tuple(x) for x in g["skl"]])
    import refio
    if os.environ.get("C5B_CPU") == "1" and refio.available("f"):     # the reference itself on one host core (about a minute)
        fa = "/tmp/prrn_c5b_%d.fa" % os.getpid()
        gen_synth.write_fasta(fa, dna)
        r = refio.run("align", fa, flavour="f", molc="n", crs=1, sh=-50, mtx="pam")
        os.unlink(fa)
        out["cpu_reference"] = {"cores": 1, "seconds": r["time"], "gcups": cells / r["time"] / 1e9, "score": r["aligns"][(0, 1)]["score"]}
    print(json.dumps(out))


def c4(ctx, members=60, length=2000, pairs=6):
    import refio
    if not refio.available("d"):
        print(json.dumps({"config": "c4", "unavailable": "oracle/_ref/ref_driver_d not built"}))
        return
    fam = gen_msa.synth_msa(members, length, 0.05, 0.35, 3, dna=True)
    staged, dumps = [], []
    tmp = "/tmp/prrn_cfg_%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    for k in range(pairs):
        cut = members * (k + 1) // (pairs + 1)
        A, B = gen_msa.split_family(fam, range(cut), range(cut, members))
        gen_msa.write_native(tmp + "/A", A, "A")
        gen_msa.write_native(tmp + "/B", B, "B")
        d = refio.run_galign(tmp + "/A", tmp + "/B", flavour="d", molc="n", ls=3, wt=1, sh=-60)
        pm, pc, h = d["pwdm"], d["pwdc"], d["header"]
        SA, SB = G.stage_pair(d["groups"][0], d["groups"][1], pm["a_mode"], pm["b_mode"], d["matrix"], dxd=(pm["DvsP"] == 0))
        gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), SA["vec"].shape[1], float(h["u"]),
                                float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
        staged.append((SA, SB, gp))
        dumps.append(d)
    t0 = time.perf_counter()
    scores, pts = ctx.align_groups(staged)
    dt = time.perf_counter() - t0
    bad = 0
    for k, d in enumerate(dumps):
        w = d["alignc"]
        bad += abs(scores[k] - w["score"]) > 1e-5 * max(1.0, abs(w["score"])) or pts[k].tolist() != w["skl"]
    cells = sum(P.group_cells(s[0], s[1], s[2].sh) for s in staged)
    cpu = sum(d["time"] for d in dumps)
    print(json.dumps({"config": "c4", "pairs": pairs, "members": members, "columns": [len(fam[0])], "Noll": dumps[0]["pwdm"]["Noll"],
                      "alnmodes": sorted(set(d["pwdm"]["alnmode"] for d in dumps)), "cells": int(cells), "seconds_e2e": dt,
                      "gcups_e2e": cells / dt / 1e9, "kernel_ms": ctx.last_kernel_ms(), "reference_alignc_seconds": cpu,
                      "reference_gcups_1core": cells / cpu / 1e9, "parity_mismatches": int(bad)}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c5a", "c5b", "c4"]
    ctx = P.Context(0)
    for w in which:
        {"c5a": c5a, "c5b": c5b, "c4": c4}[w](ctx)
    ctx.close()
