"""End to end on the GPU box: the reference's own `prrn5` program, unmodified, linked with shim/*.cc on
libprrn_gpu.so (oracle/_ref/prrn5_gpu, built where /root/reference exists and shipped with the snapshot) must print
the very MSA the plain CPU build printed -- frozen as md5 digests in tests/golden/prrn_msa.json by
`python tools/run_prrn.py --arm cpu --freeze ...` in the container that has the reference.  Every alnScoreD,
calcdist, alignC / HomScoreC and alignB_ng of the guide tree, the progressive stage and the doubly nested refinement
runs in the library; a single flipped decision anywhere changes the digest."""
import json
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_prrn  # noqa: E402

pytestmark = pytest.mark.gpu


def _run(name, env_extra=None):
    if not os.path.exists(os.path.join(run_prrn.REFDIR, run_prrn.BIN["gpu"])):
        pytest.skip("oracle/_ref/prrn5_gpu is not built (needs /root/reference at build time)")
    frozen = json.load(open(run_prrn.FROZEN))
    c = run_prrn.CONFIGS[name]
    path = run_prrn.input_path(name)
    assert path is not None
    env = {"PRRN_GPU_STATS": "1"}
    env.update(env_extra or {})
    r, _ = run_prrn.run("gpu", path, c["args"], env, timeout=900)
    assert r["rc"] == 0, r.get("stderr")
    assert r["msa_md5"] == frozen[name]["msa_md5"], "the shimmed prrn5 printed another MSA than the reference's CPU build"
    stats = " ".join(r.get("stats", []))
    assert "0 calls left on the reference's Fwd2c" in stats       # nothing ran on the CPU
    return stats


def test_prrn5_protein_small_identical_msa():
    """40 x ~200 aa, prrn5 -m blosum62: guide tree (calcdist on K1P), progressive alignment and refinement (K4 + K3)."""
    _run("small")


def test_prrn5_dna_two_piece_identical_msa():
    """BASELINE config 4 at parity size: the first 20 of the 100 DNA sequences of ~2 kb, prrn5 -yl3 (two-piece gap
    penalties, Noll = 3): 773 group alignments, hetero up to 7, uninitialised matrix entries of unused codes."""
    _run("c4n20")


def test_prrn5_best_of_n_workers_batched_identical_msa():
    """BASELINE config 3 refined from a pre-aligned start with B = 4 candidate partitions per cycle
    (prrn5 -t4 -r4: Prrn::best_of_n, src/prrn5.cc:594): the four pthread workers of a cycle reach the GPU as ONE
    pg_align_groups call through the rendezvous of shim_alignc.cc, and the MSA equals the CPU build's."""
    import re
    stats = _run("c3t4")
    m = re.search(r"largest (\d+)", stats)
    assert m and int(m.group(1)) == 4, stats
    m = re.search(r"mean batch ([0-9.]+)", stats)
    assert m and float(m.group(1)) > 1.5, stats
