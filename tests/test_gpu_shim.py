"""The drop-in boundary, end to end: the UNMODIFIED reference (oracle/_ref objects) with
shim/shim_fwd2d1.cc linked in front of its fwd2d1.o, so that the reference's own alnscore2dist /
dpscore / calcdist code calls libprrn_gpu.so for every alnScoreD.  Its output must equal the goldens
frozen from the plain reference build, byte for byte (scores, ends, distance vectors)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth  # noqa: E402
import refio  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = ["score_p24_blosum62", "score_p24_pam_f64", "score_p24_pam_f32", "score_p24_tgapf05", "score_p24_lcl15",
         "score_p24_lcl16", "score_p24_lcl15_pam_f64", "score_ragged"]


@pytest.mark.parametrize("name", CASES)
def test_reference_driver_on_gpu_library(name, tmp_path):
    g = golden(name)
    fl = g["flavour"]
    if not os.path.exists(refio.driver(fl, gpu=True)):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built (needs /root/reference at build time)" % fl)
    fa = str(tmp_path / "in.fa")
    gen_synth.write_fasta(fa, g["seqs"])
    n = len(g["seqs"])
    sc = refio.run("scores", fa, flavour=fl, gpu=True, **g["args"])
    got = [sc["scores"][(i, j)] for j in range(1, n) for i in range(j)]
    assert got == g["scores"], "alnScoreD through the shim differs from the reference"
    if "ends" in g:
        assert [list(sc["ends"][(i, j)]) for j in range(1, n) for i in range(j)] == g["ends"]
    if "dist" in g:
        ds = refio.run("dist", fa, flavour=fl, gpu=True, **g["args"])
        assert np.array_equal(ds["dist"], np.array(g["dist"])), "calcdist through the shim differs from the reference"


def test_calcdist_batch_shim_is_one_library_call(tmp_path):
    """shim/shim_calcdist.cc: the reference's calcdist(mSeq**, nn, DynScr) symbol itself is bound to the library, so
    the driver's all-vs-all step is ONE pg_calcdist call (not nn(nn-1)/2 alnScoreD calls) and still returns the
    reference's distance vector bit for bit; a mode the library refuses (lcl & 16) runs the reference's own
    calcdist (calcdist_ref) and matches its golden too."""
    import subprocess
    g = golden("score_p24_blosum62")
    fl = g["flavour"]
    drv = refio.driver(fl, gpu=True)
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % fl)
    fa = str(tmp_path / "in.fa")
    gen_synth.write_fasta(fa, g["seqs"])
    env = dict(os.environ, ALN_TAB=os.path.join(refio.REFDIR, "table"), PRRN_GPU_STATS="1")
    out = subprocess.run([drv, "dist", fa] + ["%s=%s" % kv for kv in g["args"].items()], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    n = len(g["seqs"])
    assert "prrn_gpu calcdist: %d sequences, %d pairs in one pg_calcdist call" % (n, n * (n - 1) // 2) in out.stderr
    assert np.array_equal(refio.parse(out.stdout)["dist"], np.array(g["dist"]))


ALIGN_CASES = ["align_p16_blosum62", "align_p16_pam_f32", "align_p16_twopiece_f64", "align_ragged", "align_c1_ce13a"]


@pytest.mark.parametrize("name", ALIGN_CASES)
def test_reference_align2_on_gpu_library(name, tmp_path):
    """align2 of the unmodified reference (PwdM, alignC dispatch, stdskl, the sh = -100 retry) with
    shim/shim_alignc.cc's alignC<DPunit> underneath: two single sequences."""
    g = golden(name)
    fl = g["flavour"]
    if not os.path.exists(refio.driver(fl, gpu=True)):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % fl)
    fa = str(tmp_path / "in.fa")
    gen_synth.write_fasta(fa, g["seqs"])
    al = refio.run("align", fa, flavour=fl, gpu=True, **g["args"])
    ties = 0
    for p in g["pairs"]:
        r = al["aligns"][(p["i"], p["j"])]
        assert abs(r["score"] - p["score"]) <= 1e-5 * max(1.0, abs(p["score"])), (p["i"], p["j"])
        if [list(x) for x in r["skl"]["pts"]] != [list(x) for x in p["skl"]]:
            ties += 1       # float flavour, non-integral scores: co-optimal paths within the tolerance (test_gpu_align)
            assert fl == "f"
    assert ties <= 0.05 * len(g["pairs"])


GALIGN_CASES = ["galign_c1_multi_ab_f64", "galign_gpf_prof12_raw5_wt", "galign_gpf_prof34_prof30", "galign_hlf_prof10_single",
                "galign_rhf_single_prof10", "galign_gpf_twopiece", "galign_gpf_highhetero", "galign_ngp_gapless4x3",
                "galign_gpf_lcl15", "galign_hlf_lcl10", "galign_ntv_2x2", "galign_ntv_3x1_wt", "galign_ntv_1x3_twopiece"]


@pytest.mark.parametrize("name", GALIGN_CASES)
def test_reference_group_align2_on_gpu_library(name, tmp_path):
    """Two GROUPS through the unmodified reference's PwdM staging + align2, with alignC<DPunit_hf /
    DPunit_pf> from shim/shim_alignc.cc (C++ staging through mSeqItr -> pg_align_groups -> K4 + K3)."""
    g = golden(name)
    fl = g["flavour"]
    if not os.path.exists(refio.driver(fl, gpu=True)):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % fl)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_msa

    # the goldens keep the staged view, not the MSA files: regenerate the inputs exactly as make_golden did
    import make_golden as MG
    captured = {}

    def fake_case(nm, rows_a, rows_b, flavour="d", files=None, **kv):
        captured[nm] = (rows_a, rows_b, flavour, files, kv)
    real = MG.galign_case
    MG.galign_case = fake_case
    try:
        MG.galign_cases()
    finally:
        MG.galign_case = real
    rows_a, rows_b, flavour, files, kv = captured[name]
    if files is None:
        fa, fb = str(tmp_path / "A"), str(tmp_path / "B")
        gen_msa.write_native(fa, rows_a, "A")
        gen_msa.write_native(fb, rows_b, "B")
    else:
        pytest.skip("sample files live under /root/reference only")
    env = dict(os.environ, ALN_TAB=os.path.join(refio.REFDIR, "table"), PRRN_GPU_STATS="1")
    env.pop("PRRN_GPU_ALLOW_REF", None)
    import subprocess
    out = subprocess.run([refio.driver(flavour, gpu=True), "galign", fa, "fb=" + fb] + ["%s=%s" % kvp for kvp in kv.items()],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    # the driver's alignC, HomScore and align2 calls all ran in the library (HomScore through the shim's own dispatch:
    # maln2.o folds HomScoreC into HomScore), none on the reference's Fwd2c
    assert "3 calls on the GPU (1 of them score-only" in out.stderr and "0 calls left on the reference" in out.stderr, out.stderr[-600:]
    d = refio.parse_galign(out.stdout)
    for key in ("alignc", "align2"):
        assert abs(d[key]["score"] - g[key]["score"]) <= 1e-5 * max(1.0, abs(g[key]["score"])), key
        assert d[key]["skl"] == g[key]["skl"], key
    # HomScore -> HomScoreC<recd_t> (shim_alignc.cc -> pg_score_groups): score and rr[2] of the reference
    assert abs(d["homscore"]["score"] - g["homscore"]["score"]) <= 1e-5 * max(1.0, abs(g["homscore"]["score"]))
    assert d["homscore"]["rr"] == g["homscore"]["rr"]


def _galign_inputs(name, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_msa
    import make_golden as MG
    captured = {}
    real = MG.galign_case
    MG.galign_case = lambda nm, a, b, flavour="d", files=None, **kv: captured.__setitem__(nm, (a, b, flavour, files, kv))
    try:
        MG.galign_cases()
        MG.galign_rect_cases()
        MG.galign_swg_cases()
    finally:
        MG.galign_case = real
    rows_a, rows_b, flavour, files, kv = captured[name]
    fa, fb = str(tmp_path / "A"), str(tmp_path / "B")
    gen_msa.write_native(fa, rows_a, "A")
    gen_msa.write_native(fb, rows_b, "B")
    return fa, fb, flavour, kv


RECT_CASES = ["galign_rect_ngp_gapless4x3", "galign_rect_ngp_gapless4x3_twopiece", "galign_rect_ngp_gapless3x4_wt_f32",
              "galign_rect_single_p01", "galign_rect_single_p24_twopiece", "galign_rect_single_rag03_twopiece_u1",
              "galign_rect_single_rag62"]


@pytest.mark.parametrize("name", RECT_CASES)
def test_reference_rectangle_alignC_on_gpu_library(name, tmp_path):
    """algmode.bnd = 0 on groups without gap profile: align2 calls alignC<DPunit>(seqs, pwd, scr, true) (src/maln2.cc:
    1906), i.e. forwardA + initA; the shim stages b one column further and K3 runs its rectangle form.  The driver's
    HomScore call (the rectangle form of HomScoreC, not built) is announced and left on the reference's code."""
    import subprocess
    g = golden(name)
    fa, fb, flavour, kv = _galign_inputs(name, tmp_path)
    drv = refio.driver(flavour, gpu=True)
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % flavour)
    env = dict(os.environ, ALN_TAB=os.path.join(refio.REFDIR, "table"), PRRN_GPU_ALLOW_REF="1", PRRN_GPU_STATS="1")
    out = subprocess.run([drv, "galign", fa, "fb=" + fb] + ["%s=%s" % kvp for kvp in kv.items()],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    left = [ln for ln in out.stderr.splitlines() if "left on the reference's own code" in ln]
    assert left and all("HomScoreC" in ln for ln in left), out.stderr[-800:]
    d = refio.parse_galign(out.stdout)
    assert d["pwdm"]["alnmode"] == 1
    for key in ("alignc", "align2"):
        assert abs(d[key]["score"] - g[key]["score"]) <= 1e-5 * max(1.0, abs(g[key]["score"])), key
        assert d[key]["skl"] == g[key]["skl"], key


SWG_CASES = ["galign_swg_gpf_prof12_raw5_wt", "galign_swg_gpf_twopiece", "galign_swg_hlf_prof10_single",
             "galign_swg_rhf_single_prof10_f32", "galign_swg_gpf_highhetero",
             "galign_swg_ngp_gapless4x3", "galign_swg_single_unrelated", "galign_swg_single_rag03_twopiece"]


@pytest.mark.parametrize("name", SWG_CASES)
def test_reference_smith_waterman_on_gpu_library(name, tmp_path):
    """algmode.lcl = 16: the reference's swg1st (first pass, Fwd2c::forwardC -> pg_local_groups through the shim's
    swg1st) and swg2nd (align2 inside the colony's box -> alignC on the library) must return the plain build's colony,
    score and corner list; nothing may be left on the reference's Fwd2c."""
    import subprocess
    g = golden(name)
    fa, fb, flavour, kv = _galign_inputs(name, tmp_path)
    drv = refio.driver(flavour, gpu=True)
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % flavour)
    env = dict(os.environ, ALN_TAB=os.path.join(refio.REFDIR, "table"), PRRN_GPU_STATS="1")
    env.pop("PRRN_GPU_ALLOW_REF", None)
    out = subprocess.run([drv, "galign", fa, "fb=" + fb] + ["%s=%s" % kvp for kvp in kv.items()],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    assert "1 Smith-Waterman first passes on the GPU" in out.stderr and "0 calls left on the reference" in out.stderr, out.stderr[-600:]
    d = refio.parse_galign(out.stdout)
    tol = lambda w: 1e-5 * max(1.0, abs(w))     # noqa: E731
    assert abs(d["swg"]["val"] - g["swg"]["val"]) <= tol(g["swg"]["val"])
    assert {k: d["swg"][k] for k in ("size", "mlb", "nlb", "mrb", "nrb", "lwr", "upr")} == \
           {k: g["swg"][k] for k in ("size", "mlb", "nlb", "mrb", "nrb", "lwr", "upr")}
    assert abs(d["swg2nd"]["score"] - g["swg2nd"]["score"]) <= tol(g["swg2nd"]["score"])
    assert d["swg2nd"]["skl"] == g["swg2nd"]["skl"]
    for key in ("alignc", "align2"):
        assert abs(d[key]["score"] - g[key]["score"]) <= tol(g[key]["score"]), key
        assert d[key]["skl"] == g[key]["skl"], key


def test_concurrent_workers_become_one_batch(tmp_path):
    """The rendezvous of shim_alignc.cc: B pthread workers of the reference reach alignC at the same time (as the
    workers of Prrn::best_of_n do, src/prrn5.cc:565-612) and leave with their own results from ONE pg_align_groups
    call; contexts come from the pool (no context per thread)."""
    import subprocess
    name = "galign_gpf_prof12_raw5_wt"
    g = golden(name)
    fa, fb, flavour, kv = _galign_inputs(name, tmp_path)
    drv = refio.driver(flavour, gpu=True)
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % flavour)
    env = dict(os.environ, ALN_TAB=os.path.join(refio.REFDIR, "table"), PRRN_GPU_STATS="1", PRRN_GPU_BATCH_WAIT_US="2000000")
    out = subprocess.run([drv, "galign", fa, "fb=" + fb, "mt=6"] + ["%s=%s" % kvp for kvp in kv.items()],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    d = refio.parse_galign(out.stdout)
    assert len(d["mt"]) == 6
    for m in d["mt"]:
        assert abs(m["score"] - g["align2"]["score"]) <= 1e-5 * max(1.0, abs(g["align2"]["score"]))
        assert m["skl"] == g["align2"]["skl"]
    assert "largest 6" in out.stderr, out.stderr[-400:]
    import re
    assert int(re.search(r"(\d+) contexts", out.stderr).group(1)) <= 2, out.stderr[-400:]
    # and with the rendezvous off: six library calls, same results
    env["PRRN_GPU_BATCH"] = "0"
    out = subprocess.run([drv, "galign", fa, "fb=" + fb, "mt=6"] + ["%s=%s" % kvp for kvp in kv.items()],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    assert "largest 1" in out.stderr
    assert [m["skl"] for m in refio.parse_galign(out.stdout)["mt"]] == [g["align2"]["skl"]] * 6


def test_refused_calls_are_fatal_unless_allowed(tmp_path):
    """No silent CPU fallback in the shims: a rectangle (-A) alignment is refused with the reference's fatal()
    unless PRRN_GPU_ALLOW_REF=1, and then it is announced."""
    import subprocess
    fa, fb, flavour, kv = _galign_inputs("galign_gpf_raw3x3", tmp_path)
    drv = refio.driver(flavour, gpu=True)
    if not os.path.exists(drv):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % flavour)
    env = dict(os.environ, ALN_TAB=os.path.join(refio.REFDIR, "table"))
    env.pop("PRRN_GPU_ALLOW_REF", None)
    cmd = [drv, "galign", fa, "fb=" + fb, "bnd=0"] + ["%s=%s" % kvp for kvp in kv.items()]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
    env["PRRN_GPU_ALLOW_REF"] = "1"
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "left on the reference's own code" in out.stderr


ALIGNB_CASES = ["alignb_p12_blosum62", "alignb_p12_pam_f64", "alignb_p12_twopiece_f64", "alignb_rag10_tgapf05_f64",
                "alignb_rag10_lcl15_f64", "alignb_rag10_lcl6", "alignb_rag10_lcl9_twopiece_f64", "alignb_long700"]


@pytest.mark.parametrize("name", ALIGNB_CASES)
def test_reference_alignB_ng_on_gpu_library(name, tmp_path):
    """shim/shim_fwd2b1.cc: the reference's alignB_ng / HomScoreB_ng symbols (Aln2b1, what prrn5's DynAln distances
    call, src/adjmat.cc:78-88) bound to pg_align_pairs_ng; stdskl stays the reference's own."""
    g = golden(name)
    fl = g["flavour"]
    if not os.path.exists(refio.driver(fl, gpu=True)):
        pytest.skip("oracle/_ref/ref_driver_%s_gpu is not built" % fl)
    fa = str(tmp_path / "in.fa")
    gen_synth.write_fasta(fa, g["seqs"])
    al = refio.run("alignb", fa, flavour=fl, gpu=True, **g["args"])
    exact = fl == "d" or np.all(np.nan_to_num(np.array(g["matrix"])) == np.rint(np.nan_to_num(np.array(g["matrix"]))))
    for p in g["pairs"]:
        r = al["alignb"][(p["i"], p["j"])]
        if exact:
            assert r["score"] == p["score"] and r["hom"] == p["hom"], (p["i"], p["j"])
            assert [list(x) for x in r["skl"]] == [list(x) for x in p["skl"]], (p["i"], p["j"])
        else:
            assert abs(r["score"] - p["score"]) <= 1e-5 * max(1.0, abs(p["score"]))
