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
