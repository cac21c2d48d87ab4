#!/usr/bin/env python
"""bench.py -- headline benchmark of the DP forward fill (BASELINE.json config 2).

One "step" = one all-vs-all `calcdist(DynScr)` pass (reference src/phyl.cc:318) over the synthetic
protein set: score-only banded affine Gotoh fill for every pair + the distance epilogue.

  python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (CUDA, sm_100a)
  python bench.py --impl reference --steps 2 --warmup 1    # the reference's own CPU code on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...        # the same pairs sharded over N GPUs + NCCL all-gather

Prints ONE JSON line (rank 0).  `value` = GCUPS over the cells the reference's loops visit
(SURVEY.md 8(d)), inputs resident in HBM; `e2e` = the same through the host-buffer C-ABI call
(H2D of sequences + matrix, D2H of the distance vector inside the timed region).  Further objects on the
same line: `scale_out` (config 5a, 10,000 x 300 aa, strong-scaled at the same N), `group_to_group` (kernel
K3/K4 throughput with the reference's alignC beside it), `prrn_msa` (config 3 end to end: the reference's own
prrn5, plain on the host and linked with the shims, wall seconds and MSA identity) -- the last two at N = 1.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

import gen_synth  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402

BASE_N = 1000               # config 2: 1,000 proteins of ~400 aa
SH = -60                    # prrn5 default band shoulder (src/prrn5.cc:1272)
OPS_PER_CELL = 9            # 5 add + 4 max per cell, affine score-only
DPX_OPS_PER_INSTR = 2       # one viaddmax / vimax3 retires two of those scalar operations
TRAFFIC_NCU = 1835520       # dram__bytes_read + write per k1p_score_kernel launch (profiles/r2_k1p_score_ncu_full_summary.csv)
CPU_SAMPLE_N = 400          # bounded CPU sample: first 400 sequences (79,800 pairs, ~1.1e10 cells)


def blosum62_matrix():
    with open(os.path.join(ROOT, "tests", "golden", "score_p24_blosum62.json")) as f:
        return np.array(json.load(f)["matrix"])


def workload_c2():
    """BASELINE config 2: 1,000 proteins of ~400 aa (seed 1) -- the same set at every N (strong scaling)."""
    return gen_synth.synth_set(BASE_N, 400, 0.1, 0.6, 1)


class ClockSampler:
    """SM clock / throttle-reason sampler running during the timed region: NVML (what nvidia-smi reads) polled
    every few ms from a thread, so that a timed region of tens of ms still gets samples; falls back to
    `nvidia-smi -lms` when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        self.index = index
        self.rows = []          # nvidia-smi fallback
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.nv = None
        self.h = None
        self.stop_flag = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def sample_now(self):
        if not self.nv:
            return
        nv, h = self.nv, self.h
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self.BITS:
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag:
            self.sample_now()
            time.sleep(0.004)

    def start(self):
        try:
            self.nv, self.h = self._nvml_handle()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def cpu_reference_run(seqs, threads, reps=1):
    """Time the reference's own calcdist (oracle/_ref/ref_driver_d, unmodified reference objects) --
    or the oracle port when the reference build is absent -- on the host cores.  Returns
    (gcups, kind, cores, seconds, cells).  Cells are counted in numpy (prrn_aln_b200/sharding.py: no CUDA
    library is loaded by this leg)."""
    from prrn_aln_b200 import sharding
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    cells = int(sharding.row_costs(lens, SH)[0].sum())
    import refio
    if refio.available("d"):
        fa = "/tmp/prrn_bench_cpu_%d.fa" % os.getpid()
        gen_synth.write_fasta(fa, seqs)
        r = refio.run("dist", fa, flavour="d", threads=threads, sh=SH, quiet=1, rep=reps)
        os.unlink(fa)
        return cells / r["time"] / 1e9, "reference", max(threads, 1), r["time"], cells
    import oracle_py as O
    enc = [seqcode.encode_protein(s) for s in seqs]
    M = blosum62_matrix()
    t0 = time.time()
    O.calcdist([O.seq(e) for e in enc], M, O.params(sh=SH, vtype=1), want_scores=False)
    dt = time.time() - t0
    return cells / dt / 1e9, "port", 1, dt, cells


def default_scoring_measurement(P, ctx, enc):
    """C2 again with the reference's DEFAULT scoring (PAM250 log-odds, non-integral, u=2 v=9): the exact-integer DPX
    kernels do not apply, the fill runs the IEEE float / double kernel K1F, bit-identical to the reference's VTYPE
    arithmetic.  Host-buffer calls (end to end), best of 3 after a warm-up."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "score_p24_pam_f64.json")) as f:
            M = np.array(json.load(f)["matrix"])
        ss = P.SeqSet(enc)
        out = {"metric": "DP GCUPS (calcdist, band cells), end to end", "unit": "GCUPS",
               "config": {"workload": "C2 with the reference's default PAM250 matrix (non-integral): kernel K1F"}}
        for label, vt in (("float", 0), ("double", 1)):
            prm = P.Params(P.ALPRM(sh=-60), vtype=vt)
            cells = P.calcdist_cells(ss, prm)
            ctx.calcdist(ss, prm, M)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                d = ctx.calcdist(ss, prm, M)
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            out[label] = {"value": cells / best / 1e9, "call_ms": best * 1e3, "checksum": float(np.sum(d, dtype=np.float64))}
        return out
    except Exception as e:
        return {"error": repr(e)[:300]}


def long_pair_measurement(P, ctx):
    """BASELINE config 5, second half: ONE DNA pair of 30 kb x 30 kb -- alignC<DPunit> with path on the striped
    wavefront kernel of K2 (single-warp stripes of 128 rows, bottom rows handed down through flag-less slots in L2),
    device traceback included.  Host-buffer calls, best of 5 after the first call (which grows the direction-bit
    store); the corner list is compared with the reference's own `aln` run on this pair (tests/golden)."""
    try:
        from prrn_aln_b200 import seqcode
        dna = gen_synth.synth_set(2, 30000, 0.2, 0.2, 5, gen_synth.NT)
        e2 = [seqcode.encode_dna(s) for s in dna]
        Mn = np.full((18, 18), -4.0)
        np.fill_diagonal(Mn, 2.0)
        prm = P.Params(P.ALPRM(u=2, v=6, sh=-50))
        ss = P.SeqSet(e2)
        cells = int(P.calcdist_cells(ss, prm))
        ctx.align_pairs(ss, [0], [1], prm, Mn)
        best, fill = None, None
        for _ in range(5):
            t0 = time.perf_counter()
            sc, raw = ctx.align_pairs(ss, [0], [1], prm, Mn)
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, fill = dt, ctx.last_kernel_ms()
        pts = P.stdskl(raw[0])
        out = {"metric": "DP GCUPS (one pair with path, band cells)", "unit": "GCUPS", "value": cells / (fill * 1e-3) / 1e9,
               "fill_kernel_ms": fill, "e2e": {"value": cells / best / 1e9, "unit": "GCUPS", "call_ms": best * 1e3},
               "cells": cells, "score": float(sc[0]), "corners": len(pts), "replicas_only": True,
               "config": {"workload": "C5b: one synthetic DNA pair 29,979 x 30,017 nt (seed 5, 20 % divergence), match 2 "
                                      "mismatch -4, u=2 v=6 sh=-50; fill + device traceback + D2H of the corner list"}}
        gpath = os.path.join(ROOT, "tests", "golden", "align_c5b_30k.json")
        if os.path.exists(gpath):
            with open(gpath) as f:
                g = json.load(f)["pairs"][0]
            out["equals_reference_alignment"] = bool(float(sc[0]) == g["score"] and pts == [tuple(x) for x in g["skl"]])
        return out
    except Exception as e:
        return {"error": repr(e)[:300]}


def group_side_measurement():
    """The other half of BASELINE.json's metric: group-to-group DP (kernel K3 + K4) on partitions of a
    200 x ~500 aa family (config 3 shape), with the reference's alignC timed beside it on one host core.
    The inputs are staged by the reference itself (oracle/_ref/ref_driver_d: PwdM -> mkthick / Gfq /
    convseq), which is test infrastructure; without it this object only says so."""
    try:
        import argparse as _ap
        import bench_groups
        import refio
        if not refio.available("d"):
            return {"unavailable": "oracle/_ref/ref_driver_d (the reference's staging of groups) is not built"}
        a = _ap.Namespace(members=200, length=500, pairs=24, sh=-60, seed=7, cpu_rep=1)
        dumps = bench_groups.build_pairs(a)
        return bench_groups.measure(dumps, replicate=16, steps=3)
    except Exception as e:      # the headline line must survive a failure of the side measurement
        return {"error": repr(e)[:300]}


def group_candidates_measurement(P, torch, dist, rank, world, local):
    """the B candidate partitions of one refinement step (Prrn::best_of_n, src/prrn5.cc:594) are
    independent -- B / N per GPU by a longest-processing-time split of their DP cells (sharding.shard_candidates), every
    rank aligns its share with pg_align_groups, ONE all-reduce(MAX) of B doubles over NCCL picks the winner."""
    try:
        import argparse as _ap
        import bench_groups
        import refio
        from prrn_aln_b200 import groups as G
        from prrn_aln_b200 import sharding
        if not refio.available("d"):
            return {"unavailable": "oracle/_ref/ref_driver_d (the reference's staging of groups) is not built"}
        a = _ap.Namespace(members=200, length=500, pairs=24, sh=-60, seed=7, cpu_rep=1)
        dumps = bench_groups.build_pairs(a)             # every rank stages the same candidates (replicated MSA state)
        rep = 8
        staged, costs = [], []
        for d in dumps:
            pm, pc, h = d["pwdm"], d["pwdc"], d["header"]
            A, B = G.stage_pair(d["groups"][0], d["groups"][1], pm["a_mode"], pm["b_mode"], d["matrix"], dxd=(pm["DvsP"] == 0))
            gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                    float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
            staged.append((A, B, gp))
            costs.append(int(P.group_cells(A, B, gp.sh)))
        staged, costs = staged * rep, costs * rep
        ctx = P.Context(local)

        def scores_of(idx):
            return list(ctx.align_groups([staged[i] for i in idx])[0])
        for _ in range(3):
            sharding.best_of_n_sharded(scores_of, costs, rank, world, dist)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        steps = 3
        t0 = time.perf_counter()
        for _ in range(steps):
            best, val, allv = sharding.best_of_n_sharded(scores_of, costs, rank, world, dist)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ctx.close()
        want = [d["alignc"]["score"] for d in dumps] * rep
        bad = sum(1 for x, w in zip(allv.tolist(), want) if abs(x - w) > 1e-5 * max(1.0, abs(w)))
        return {"metric": "group-to-group DP GCUPS (candidate partitions sharded over GPUs, host-buffer calls)",
                "value": sum(costs) * steps / float(dt.item()) / 1e9, "unit": "GCUPS", "n_gpus": world, "ms_per_step": 1e3 * float(dt.item()) / steps,
                "candidates": len(costs), "cells": int(sum(costs)), "best": best, "best_score": val, "score_mismatches_vs_reference": bad,
                "collective": "nccl all_reduce(max) of %d doubles" % len(costs) if world > 1 else None, "scaling": "strong"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def edge_candidates_measurement(P, torch, dist, rank, world, local):
    """candidate scoring of the sparse distance graph (AdjacentMat::spaln_job, src/adjmat.cc:
    119-156, DynScr branch): per query the k-mer search returns a few database sequences and every (query, candidate)
    pair gets 100 * alnscore2dist.  Here: the C5a set, 32 seeded candidates per query (the search itself is host code
    outside the path), queries dealt to the ranks by DP cells, pg_dist_pairs per rank (host-buffer call), ONE
    all-reduce(SUM) of the edge vector."""
    try:
        from prrn_aln_b200 import seqcode, sharding
        seqs = gen_synth.config_set("c5a")
        enc = [seqcode.encode_protein(x) for x in seqs]
        ss = P.SeqSet(enc)
        n = len(enc)
        rng = np.random.default_rng(11)
        per = 32
        qi = np.repeat(np.arange(n), per)
        si = (qi + 1 + rng.integers(0, n - 1, size=len(qi))) % n       # never the query itself
        M = blosum62_matrix()
        prm = P.Params(P.ALPRM(sh=-60), vtype=1)
        lens = np.array([len(e) for e in enc])
        ctx = P.Context(local)

        def dist_of(idx):
            return ctx.dist_pairs(ss, qi[idx], si[idx], prm, M)
        for _ in range(2):
            sharding.dist_edges_sharded(dist_of, qi, si, lens, -60, rank, world, dist)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        steps = 3
        t0 = time.perf_counter()
        for _ in range(steps):
            full, mine = sharding.dist_edges_sharded(dist_of, qi, si, lens, -60, rank, world, dist)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # parity: a seeded sample of this rank's edges against the all-pairs entry of the same pair (a = query, b = hit)
        bad = 0
        if len(mine):
            pick = mine[rng.integers(0, len(mine), size=64)]
            for e in pick:
                a_, b_ = int(qi[e]), int(si[e])
                ref = ctx.dist_pairs(P.SeqSet([enc[a_], enc[b_]]), [0], [1], prm, M)[0]
                bad += float(full[e]) != float(ref)
        ctx.close()
        cells = int(sharding.pair_costs(lens, qi, si, -60).sum())
        return {"metric": "DP GCUPS (candidate edges of the sparse distance graph, sharded by query, host-buffer calls)",
                "value": cells * steps / float(dt.item()) / 1e9, "unit": "GCUPS", "n_gpus": world,
                "ms_per_step": 1e3 * float(dt.item()) / steps, "edges": int(len(qi)), "cells": cells,
                "checksum": float(full.sum().item()), "sample_mismatches": int(bad),
                "collective": "nccl all_reduce(sum) of %d doubles" % len(qi) if world > 1 else None, "scaling": "strong"}
    except Exception as e:
        return {"error": repr(e)[:300]}


def prrn_msa_measurement():
    """The third part of BASELINE.json's metric: `prrn` MSA wall seconds on config 3 (200 x ~500 aa, prrn5 -m blosum62):
    the reference's own prrn5 program as it is (every DP on ONE host core: upstream's threaded mode crashes on
    unaligned protein input) and the same objects linked with shim/*.cc (every alnScoreD / calcdist /
    alignC in libprrn_gpu.so), side by side in the same run; the two MSAs must be identical."""
    try:
        import run_prrn
        if not all(os.path.exists(os.path.join(run_prrn.REFDIR, run_prrn.BIN[a])) for a in ("cpu", "gpu")):
            return {"unavailable": "oracle/_ref/prrn5_cpu / prrn5_gpu (the reference's prrn5, plain and shimmed) are not built"}
        c = run_prrn.CONFIGS["c3"]
        path = run_prrn.input_path("c3")
        res = {}

        def leg(arm):
            res[arm] = run_prrn.run(arm, path, c["args"], {"PRRN_GPU_STATS": "1"} if arm == "gpu" else {})[0]
        th = [threading.Thread(target=leg, args=(a,)) for a in ("cpu", "gpu")]      # side by side: different resources
        for t in th:
            t.start()
        for t in th:
            t.join()
        out = {"metric": "prrn MSA wall seconds", "unit": "s", "higher_is_better": False,
               "config": {"workload": "C3: prrn5 -m blosum62 on 200 x ~500 aa synthetic proteins (seed 1), double VTYPE, -t0"},
               "value": res["gpu"]["wall_s"], "cpu_baseline": {"value": res["cpu"]["wall_s"], "unit": "s", "cores": 1,
                                                               "kind": "reference", "sample": "the whole run"},
               "identical_msa": res["cpu"]["rc"] == 0 and res["gpu"]["rc"] == 0 and res["cpu"]["msa_md5"] == res["gpu"]["msa_md5"],
               "msa_md5": res["gpu"]["msa_md5"], "rc": [res["cpu"]["rc"], res["gpu"]["rc"]],
               "library_stats": res["gpu"].get("stats")}
        frozen = json.load(open(run_prrn.FROZEN)) if os.path.exists(run_prrn.FROZEN) else {}
        if "c3" in frozen:
            out["equals_frozen_reference_msa"] = frozen["c3"]["msa_md5"] == res["gpu"]["msa_md5"]
        return out
    except Exception as e:
        return {"error": repr(e)[:300]}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seqs = workload_c2()[:CPU_SAMPLE_N]
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_run(seqs[:100], threads)
    t_tot, cells_tot, kind, cores = 0.0, 0, "port", 1
    for _ in range(args.steps):
        g, kind, cores, dt, cells = cpu_reference_run(seqs, threads)
        t_tot += dt
        cells_tot += cells
    val = cells_tot / t_tot / 1e9
    sample = "first %d of the 1000 sequences (%d pairs, %.3g cells) per step" % (
        CPU_SAMPLE_N, CPU_SAMPLE_N * (CPU_SAMPLE_N - 1) // 2, cells_tot / args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "DP GCUPS (all-pairs calcdist, band cells)", "value": val, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 all-vs-all calcdist(DynScr): 1000 x ~400 aa synthetic proteins (seed 1), "
                               "BLOSUM62 u=2 v=9 sh=-60; CPU arm runs a bounded sample", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


class AllPairs:
    """One all-vs-all calcdist workload, STRONG-scaled: the condensed index range of the same sequence set is cut
    into `world` contiguous, cost-balanced shards (equal DP cells, prrn_aln_b200/sharding.py); every rank fills its
    shard with pg_calcdist_dev and the shards are exchanged by NCCL all-gather.  With chunks > 1 a rank's shard is cut
    again and the all-gather of piece c runs (asynchronously, on NCCL's stream) while piece c + 1 is being filled."""

    def __init__(self, P, torch, dist, ctx, seqs, rank, world, chunks, stream):
        from prrn_aln_b200 import sharding
        self.P, self.torch, self.dist, self.ctx, self.rank, self.world, self.stream = P, torch, dist, ctx, rank, world, stream
        self.enc = [seqcode.encode_protein(s) for s in seqs]
        self.ss = P.SeqSet(self.enc)
        self.M = blosum62_matrix()
        self.prm = P.Params(P.ALPRM(sh=SH), vtype=1)     # prrn build: FTYPE = double
        self.npair = self.ss.n * (self.ss.n - 1) // 2
        lens = np.array([len(e) for e in self.enc], dtype=np.int64)
        self.cells_total = int(sharding.row_costs(lens, SH)[0].sum())
        # world * chunks cost-balanced pieces; rank r owns pieces r, r + world, ... (piece c of every rank is
        # exchanged in one all-gather)
        pieces = sharding.cost_balanced_ranges(lens, SH, world * chunks)
        self.mine = [pieces[c * world + rank] for c in range(chunks)]
        self.slot = [max(b - a for a, b in pieces[c * world:(c + 1) * world]) for c in range(chunks)]
        # one device handle per piece: each caches the host-built schedule of its own k range
        self.dseqs = [ctx.upload(self.ss) for _ in range(chunks)]
        self.d_shard = [torch.empty(max(n, 1), dtype=torch.float64, device="cuda") for n in self.slot]
        self.d_all = [torch.empty(max(n, 1) * world, dtype=torch.float64, device="cuda") if world > 1 else self.d_shard[c]
                      for c, n in enumerate(self.slot)]
        self.pieces = pieces
        self.chunks = chunks

    def step(self):
        nl = 0
        handles = []
        for c, (k0, k1) in enumerate(self.mine):
            nl += self.ctx.calcdist_dev(self.dseqs[c], self.prm, self.M, k0, k1, self.d_shard[c].data_ptr(), self.stream.cuda_stream)
            if self.world > 1:
                handles.append(self.dist.all_gather_into_tensor(self.d_all[c], self.d_shard[c], async_op=True))
        for h in handles:
            h.wait()
        return nl

    def assembled(self):
        """The full condensed distance vector on this rank (device), pieces put back in k order."""
        torch = self.torch
        parts = []
        for idx, (a, b) in enumerate(self.pieces):
            c, r = divmod(idx, self.world)
            parts.append((a, self.d_all[c][r * max(self.slot[c], 1): r * max(self.slot[c], 1) + (b - a)] if self.world > 1
                          else self.d_shard[c][:b - a]))
        parts.sort(key=lambda t: t[0])
        return torch.cat([p for _, p in parts])

    def close(self):
        for d in self.dseqs:
            self.ctx.free_seqs(d)


def time_steps(torch, dist, world, rank, stream, step, steps, warmup, flush, sampler=None):
    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        step()
    sync_all()
    launches = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync_all()
    if sampler is not None:
        sampler.start()                  # polls while the timed steps below execute
    for i in range(steps):
        flush.fill_(i & 0xff)            # evict L2 between timed iterations (outside the events)
        ev[i][0].record(stream)
        launches += step()
        ev[i][1].record(stream)
    if sampler is not None:
        sampler.sample_now()             # the steps are queued and running: at least one sample under load
    sync_all()
    clocks = sampler.stop() if sampler is not None else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), launches, clocks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-groups", action="store_true", help="skip the group-to-group (K3) side measurement")
    ap.add_argument("--no-scaleout", action="store_true", help="skip the C5a (10,000 x 300 aa) strong-scaling object")
    ap.add_argument("--no-prrn", action="store_true", help="skip the prrn MSA wall-time object (C3, ~80 s of host time)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import prrn_aln_b200 as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = P.Context(local)
    stream = torch.cuda.Stream()         # a real (non-default) stream: kernels, NCCL and timing events share it
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    # ---- headline: C2 (1,000 x ~400 aa), the same set at every N: strong scaling
    c2 = AllPairs(P, torch, dist, ctx, workload_c2(), rank, world, 1, stream)
    ms_total, launches, clocks = time_steps(torch, dist, world, rank, stream, c2.step, args.steps, args.warmup, flush,
                                            ClockSampler(local) if rank == 0 else None)
    gcups = c2.cells_total * args.steps / (ms_total * 1e-3) / 1e9
    ss, npair, M, prm = c2.ss, c2.npair, c2.M, c2.prm
    c2_enc = c2.enc
    k0, k1 = c2.mine[0]

    # ---- end to end through the host-buffer C-ABI call: H2D (sequences, matrix) + D2H (distances)
    pin_out = torch.empty(max(k1 - k0, 1), dtype=torch.float64).pin_memory()
    out_np = pin_out.numpy()
    res_pin = torch.from_numpy(ss.res.copy()).pin_memory()
    ss_pin = P.SeqSet(c2.enc)
    ss_pin.res = res_pin.numpy()
    for _ in range(3):
        ctx.calcdist(ss_pin, prm, M, k0, k1, out=out_np)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.calcdist(ss_pin, prm, M, k0, k1, out=out_np)
        checksum = float(out_np[:8].sum())  # noqa: F841
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_gcups = c2.cells_total * args.steps / float(t.item()) / 1e9
    h2d = int(ss.res.nbytes + ss.offs.nbytes + 2 * ss.lens.nbytes + ss.n + M.size * 4)
    d2h = int((k1 - k0) * 8)

    # ---- roofline of the dominant kernel (the fill): DPX issue rate measured on this device
    peak32, peak16 = ctx.dpx_peak()
    # The C2 workload fits 16 bits (k1p_fits), so the fill runs the packed int16x2 kernel: one DPX
    # instruction retires 2 scalar ops on each of 2 packed cells.  Peak = measured s16x2 issue rate x 4.
    achieved = gcups / world * OPS_PER_CELL              # per GPU, Gop/s of algorithmic add/max
    peak = peak16 * DPX_OPS_PER_INSTR * 2
    roofline = {"bound": "dpx-int16x2 issue", "achieved": achieved, "peak": peak, "unit": "Gop/s",
                "frac": achieved / peak, "traffic": TRAFFIC_NCU,
                "peak_source": "measured here: pg_dpx_peak register-only __viaddmax_s16x2 chains = %.0f G thread-instr/s "
                               "(x2 scalar ops x2 packed cells each); MEASURED_PEAKS.json has no integer peak" % peak16,
                "ops_per_cell": OPS_PER_CELL, "kernel_dpx_instr_per_cell": 2,
                "issue_frac": gcups / world * 2 / peak16, "dpx_s32_ginstr": peak32,
                "hbm_bytes_per_step": h2d + d2h}

    out = {
        "metric": "DP GCUPS (all-pairs calcdist, band cells)", "value": gcups, "unit": "GCUPS", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
        "config": {"workload": "C2 all-vs-all calcdist(DynScr): %d x ~400 aa synthetic proteins (seed 1), %d pairs, "
                               "BLOSUM62 u=2 v=9 sh=-60, prrn (double) build semantics; the same set at every N, "
                               "condensed index cut into %d cost-balanced shard(s) (equal DP cells)" % (ss.n, npair, world),
                   "pairs": npair, "cells": c2.cells_total, "full_matrix_cells": int(sum(
                       int(ss.lens[j]) * int(ss.lens[:j].sum()) for j in range(1, ss.n))),
                   "l2": "256 MB buffer written between timed iterations", "collective": "nccl all_gather" if world > 1 else None,
                   "schedule": "value: the host-built work schedule of the shard is cached on the device-sequence handle "
                               "(built once, in the warm-up); e2e: every call uploads the sequences and rebuilds it"},
        "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    c2.close()

    # ---- scale-out config C5a: 10,000 x ~300 aa (49,995,000 pairs), strong scaling, all-gather overlapped
    if not args.no_scaleout:
        try:
            chunks = 1 if world == 1 else 4
            c5 = AllPairs(P, torch, dist, ctx, gen_synth.config_set("c5a"), rank, world, chunks, stream)
            st5 = 3
            ms5, nl5, _ = time_steps(torch, dist, world, rank, stream, c5.step, st5, 3, flush)
            ag_ms = None
            if world > 1:               # the exchange alone (all pieces back to back), for the record
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); dist.barrier()
                e0.record(stream)
                for c in range(c5.chunks):
                    dist.all_gather_into_tensor(c5.d_all[c], c5.d_shard[c])
                e1.record(stream)
                torch.cuda.synchronize()
                ag_ms = e0.elapsed_time(e1)
            full = c5.assembled()
            out["scale_out"] = {
                "metric": "DP GCUPS (all-pairs calcdist, band cells)", "value": c5.cells_total * st5 / (ms5 * 1e-3) / 1e9,
                "unit": "GCUPS", "n_gpus": world, "steps": st5, "warmup": 3, "ms_per_step": ms5 / st5, "scaling": "strong",
                "gpu_launches": nl5, "all_gather_ms_alone": ag_ms, "all_gather_bytes": int(c5.npair * 8),
                "checksum": float(full.sum().item()), "pairs": c5.npair, "cells": c5.cells_total,
                "config": {"workload": "C5a all-vs-all calcdist(DynScr): 10,000 x ~300 aa synthetic proteins (seed 5), "
                                       "49,995,000 pairs, BLOSUM62 u=2 v=9 sh=-60; the same set at every N, %d cost-balanced "
                                       "piece(s) per rank, the all-gather of a piece overlaps the fill of the next" % chunks}}
            c5.close()
            del c5, full
        except Exception as e:
            out["scale_out"] = {"error": repr(e)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = workload_c2()[:CPU_SAMPLE_N]
        g, kind, cores, dtc, cells = cpu_reference_run(sample, os.cpu_count() or 1)
        out["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind,
                               "sample": "first %d of the 1000 sequences: %d pairs, %.3g cells, %.2f s" % (
                                   CPU_SAMPLE_N, CPU_SAMPLE_N * (CPU_SAMPLE_N - 1) // 2, cells, dtc)}
    if rank == 0 and world == 1 and not args.no_groups:
        out["default_scoring"] = default_scoring_measurement(P, ctx, c2_enc)
        out["long_pair"] = long_pair_measurement(P, ctx)
    ctx.close()
    if world > 1 and not args.no_groups:
        gc = group_candidates_measurement(P, torch, dist, rank, world, local)
        if rank == 0:
            out["group_candidates"] = gc
    if not args.no_groups:
        ec = edge_candidates_measurement(P, torch, dist, rank, world, local)
        if rank == 0:
            out["edge_candidates"] = ec
    if rank == 0 and world == 1 and not args.no_groups:
        out["group_to_group"] = group_side_measurement()
    if rank == 0 and world == 1 and not args.no_prrn:
        out["prrn_msa"] = prrn_msa_measurement()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
