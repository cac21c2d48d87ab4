"""Contexts, streams and threads: the cases the round-1 advisor flagged.
* pg_calcdist_dev on caller-supplied streams shares the context's workspace (work-queue counter, matrix, items,
  self scores): back-to-back calls on DIFFERENT streams, and a host-buffer call right after, must not overwrite
  each other's workspace under a running kernel (the workspace passes from stream to stream through an event).
* Two host threads with a context each run different entry points at the same time."""
import json
import os
import sys
import threading

import numpy as np
import pytest

from conftest import ROOT, golden, golden_names
import prrn_aln_b200 as P
from prrn_aln_b200 import groups as G
from prrn_aln_b200 import seqcode

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _set(n, length, seed):
    return P.SeqSet([seqcode.encode_protein(s) for s in gen_synth.synth_set(n, length, 0.1, 0.6, seed)])


def test_calcdist_dev_on_two_streams_and_a_host_call_do_not_collide():
    import torch
    M = np.array(golden("score_p24_blosum62")["matrix"])
    prm = P.Params(P.ALPRM(sh=-60), vtype=1)
    ctx = P.Context(0)
    a, b = _set(300, 300, 3), _set(260, 350, 4)
    want_a, want_b = ctx.calcdist(a, prm, M), ctx.calcdist(b, prm, M)
    da, db = ctx.upload(a), ctx.upload(b)
    na, nb = a.n * (a.n - 1) // 2, b.n * (b.n - 1) // 2
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(4):
        oa = torch.zeros(na, dtype=torch.float64, device="cuda")
        ob = torch.zeros(nb, dtype=torch.float64, device="cuda")
        ctx.calcdist_dev(da, prm, M, 0, na, oa.data_ptr(), s1.cuda_stream)      # asynchronous
        ctx.calcdist_dev(db, prm, M, 0, nb, ob.data_ptr(), s2.cuda_stream)      # another stream, same workspace
        host = ctx.calcdist(a, prm, M, 100, 5000)                               # and the library's own stream
        torch.cuda.synchronize()
        assert np.array_equal(oa.cpu().numpy(), want_a)
        assert np.array_equal(ob.cpu().numpy(), want_b)
        assert np.array_equal(host, want_a[100:5000])
    ctx.free_seqs(da); ctx.free_seqs(db)
    ctx.close()


def test_two_threads_two_contexts_different_entry_points():
    M = np.array(golden("score_p24_blosum62")["matrix"])
    prm = P.Params(P.ALPRM(sh=-60), vtype=1)
    ss = _set(200, 250, 9)
    gs = [golden(n) for n in golden_names("galign_")]

    def stage(g):
        pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
        A, B = G.stage_pair(g["groups"][0], g["groups"][1], pm["a_mode"], pm["b_mode"], g["matrix"], dxd=(pm["DvsP"] == 0))
        gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                                float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
        return A, B, gp
    staged = [stage(g) for g in gs]
    c0 = P.Context(0)
    want_d = c0.calcdist(ss, prm, M)
    c0.close()
    errs = []

    def dist_worker():
        try:
            c = P.Context(0)
            for _ in range(6):
                if not np.array_equal(c.calcdist(ss, prm, M), want_d):
                    errs.append("calcdist differs under concurrency")
            c.close()
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    def group_worker():
        try:
            c = P.Context(0)
            for _ in range(6):
                scores, pts = c.align_groups(staged)
                for k, g in enumerate(gs):
                    w = g["alignc"]
                    if abs(scores[k] - w["score"]) > 1e-5 * max(1.0, abs(w["score"])) or pts[k].tolist() != w["skl"]:
                        errs.append("group alignment %s differs under concurrency" % g["name"])
            c.close()
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=dist_worker), threading.Thread(target=group_worker)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[:3]


def test_score_groups_edge_cases():
    ctx = P.Context(0)
    s, rr = ctx.score_groups([])
    assert len(s) == 0 and rr.shape == (0, 2)
    g = golden("galign_gpf_raw3x3")
    pm, pc, h = g["pwdm"], g["pwdc"], g["header"]
    A, B = G.stage_pair(g["groups"][0], g["groups"][1], pm["a_mode"], pm["b_mode"], g["matrix"])
    gp = P.gparams_from_pwd(pm["alnmode"], pm["Noll"], pm["codonk1"], int(h["sh"]), A["vec"].shape[1], float(h["u"]),
                            float(h["v"]), pc["vgop1"], pc["BasicGOP"], pc["BasicGEP"], pc["LongGOP"], pc["LongGEP"])
    gp.alnmode = 3              # RHF_ALN (rectangle): refused, never computed some other way
    with pytest.raises(P.PgError) as e:
        ctx.score_groups([(A, B, gp)])
    assert e.value.code == 4
    ctx.close()
