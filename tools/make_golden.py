#!/usr/bin/env python
"""Freeze golden vectors by running the UNMODIFIED reference (oracle/_ref/ref_driver_{f,d}, built by
oracle/Makefile from /root/reference/src) on fixed-seed inputs.  Only runs where /root/reference was
present at build time (this container); the JSON files it writes under tests/golden/ are committed
and are what travels to the GPU box.

    python tools/make_golden.py            # (re)generate everything
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import gen_synth  # noqa: E402
import refio  # noqa: E402
from prrn_aln_b200 import seqcode  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TMP = "/tmp/prrn_golden"


def score_case(name, seqs, flavour="f", dna=False, **kv):
    """alnScoreD per pair + calcdist(DynScr) vector + the Simmtx the reference built."""
    os.makedirs(TMP, exist_ok=True)
    fa = os.path.join(TMP, name + ".fa")
    gen_synth.write_fasta(fa, seqs)
    extra = dict(kv)
    if dna:
        extra["molc"] = "n"
    sc = refio.run("scores", fa, flavour=flavour, **extra)
    mt = refio.run("matrix", fa, flavour=flavour, **extra)
    n = len(seqs)
    scores = [sc["scores"][(i, j)] for j in range(1, n) for i in range(j)]
    rec = dict(name=name, kind="score", flavour=flavour, dna=dna, params=sc["header"], args=kv,
               seqs=seqs, matrix=mt["matrix"].tolist(), scores=scores)
    lcl = int(kv.get("lcl", 0))
    if sc["ends"]:      # semi-global with end points (Fwd2d_vd): ends[2] per pair
        rec["ends"] = [list(sc["ends"][(i, j)]) for j in range(1, n) for i in range(j)]
    if not (lcl & 16):  # calcdist with lcl & 16 reads uninitialised `ends` in the reference (aln2.cc:296-305)
        rec["dist"] = refio.run("dist", fa, flavour=flavour, **extra)["dist"].tolist()
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(rec, f)
    print("wrote", name, "pairs", len(scores))


def align_case(name, seqs, flavour="f", **kv):
    """align2 per pair (alignC<DPunit> + stdskl): score and normalised corner list."""
    os.makedirs(TMP, exist_ok=True)
    fa = os.path.join(TMP, name + ".fa")
    gen_synth.write_fasta(fa, seqs)
    al = refio.run("align", fa, flavour=flavour, **kv)
    mt = refio.run("matrix", fa, flavour=flavour, **kv)
    n = len(seqs)
    pairs = []
    for j in range(1, n):
        for i in range(j):
            r = al["aligns"][(i, j)]
            assert r["swp"] == 0
            pairs.append(dict(i=i, j=j, mode=r["mode"], score=r["score"], skl=r["skl"]["pts"], fstat=al["fstat"][(i, j)]))
    rec = dict(name=name, kind="align", flavour=flavour, params=al["header"], args=kv, seqs=seqs,
               matrix=mt["matrix"].tolist(), pairs=pairs)
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(rec, f)
    print("wrote", name, "pairs", len(pairs))


def dna_pair_cases():
    """Long DNA pairs through aln's own set-up (prePwd(Seq**), algmode.crs = 1: s[=] 2, s[#] -4, u 2, v 6; sh -50):
    C5b (30 kb x 30 kb, 55 s of the reference's alignC<DPunit>) and a 6 kb pair -- the striped long-pair kernel."""
    align_case("align_dna6k", gen_synth.synth_set(2, 6000, 0.2, 0.2, 7, gen_synth.NT), molc="n", crs=1, sh=-50, mtx="pam")
    align_case("align_c5b_30k", gen_synth.synth_set(2, 30000, 0.2, 0.2, 5, gen_synth.NT), molc="n", crs=1, sh=-50, mtx="pam")


def lcl_cases(p24, rag):
    """algmode.lcl variants of alnScoreD: SWG score (lcl & 16) and semi-global with `ends` (Fwd2d_vd)."""
    score_case("score_p24_lcl16", p24, lcl=16)
    score_case("score_p24_lcl16_pam_f64", p24, flavour="d", mtx="pam", lcl=16, sh=-30)
    score_case("score_rag_lcl16", rag, lcl=16, sh=5)
    score_case("score_p24_lcl15", p24, lcl=15)
    score_case("score_p24_lcl5_tgapf05", p24, lcl=5, tgapf=0.5)
    score_case("score_p24_lcl10_pam_f32", p24, mtx="pam", lcl=10)
    score_case("score_p24_lcl15_pam_f64", p24, flavour="d", mtx="pam", lcl=15, sh=-30)
    score_case("score_rag_lcl15", rag, lcl=15, sh=5)
    score_case("score_p24_tgapf03_pam_f64", p24, flavour="d", mtx="pam", tgapf=0.3)


def galign_case(name, rows_a, rows_b, flavour="d", files=None, **kv):
    """Group-to-group alignC: the reference's staged inputs (what Fwd2c reads through mSeqItr after
    PwdM) + score + raw corner list (Vmf back-walk order) + align2's normalised list."""
    import gen_msa
    os.makedirs(TMP, exist_ok=True)
    if files is None:
        fa, fb = os.path.join(TMP, name + "_A"), os.path.join(TMP, name + "_B")
        gen_msa.write_native(fa, rows_a, "A")
        gen_msa.write_native(fb, rows_b, "B")
    else:
        fa, fb = files
    d = refio.run_galign(fa, fb, flavour=flavour, **kv)
    d.pop("time", None)
    d.update(name=name, kind="galign", flavour=flavour, args=kv)
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(d, f, separators=(",", ":"))
    print("wrote", name, "alnmode", d["pwdm"]["alnmode"], "a/b mode", d["pwdm"]["a_mode"], d["pwdm"]["b_mode"],
          "hetero", d["groups"][0]["hetero"], d["groups"][1]["hetero"], "score", d["alignc"]["score"])


def galign_cases():
    import gen_msa
    sp = "/root/reference/sample/pas/"
    galign_case("galign_c1_multi_ab_f32", None, None, flavour="f", files=(sp + "Multi_A", sp + "Multi_B"), mtx="blosum62", sh=-50)
    galign_case("galign_c1_multi_ab_f64", None, None, flavour="d", files=(sp + "Multi_A", sp + "Multi_B"), mtx="blosum62", sh=-50)
    fam = gen_msa.synth_msa(64, 90, 0.1, 0.5, 41)
    A, B = gen_msa.split_family(fam, range(0, 3), range(3, 6))
    galign_case("galign_gpf_raw3x3", A, B, mtx="blosum62")                       # GPF, sim22i
    A, B = gen_msa.split_family(fam, range(0, 12), range(12, 17))
    galign_case("galign_gpf_prof12_raw5", A, B)                                   # GPF, sim32, PAM
    galign_case("galign_gpf_prof12_raw5_wt", A, B, wt=1)                          # GPF, sim32w (weights)
    galign_case("galign_gpf_prof12_raw5_f32", A, B, flavour="f", wt=1)
    A, B = gen_msa.split_family(fam, range(0, 34), range(34, 64))
    galign_case("galign_gpf_prof34_prof30", A, B, wt=1, sh=-30)                   # GPF, sim33
    A, B = gen_msa.split_family(fam, range(0, 10), [10])
    galign_case("galign_hlf_prof10_single", A, B)                                 # HLF, sim31
    galign_case("galign_rhf_single_prof10", B, A, wt=1)                           # RHF (swap)
    A, B = gen_msa.split_family(fam, range(0, 8), range(8, 14))
    galign_case("galign_gpf_twopiece", A, B, ls=3, wt=1)                          # Noll = 3
    hh = gen_msa.synth_msa(40, 80, 0.3, 0.8, 45, indel_events=12.0)                # many distinct gap lengths per column
    A, B = gen_msa.split_family(hh, range(0, 22), range(22, 40))
    galign_case("galign_gpf_highhetero", A, B, wt=1, sh=-40)
    A, B = gen_msa.split_family(hh, range(0, 3), range(3, 6))
    galign_case("galign_gpf_raw3x3_gappy", A, B, mtx="blosum62")                 # GPF, sim22i
    # semi-global (free end gaps on some sides): exg_seq turns terminal gap runs into nil columns, the
    # thickness / gap-profile staging carries the effect (inex.nils -> pua per cell in forwardB)
    def ragged(rows, seed):
        """members that start late / end early: terminal gap runs of 2..9 columns in a third of the rows"""
        import random
        rng = random.Random(seed)
        out = []
        for r in rows:
            r = list(r)
            if rng.random() < 0.35:
                for k in range(rng.randint(2, 9)):
                    r[k] = "-"
            if rng.random() < 0.35:
                for k in range(rng.randint(2, 9)):
                    r[-1 - k] = "-"
            out.append("".join(r))
        return out
    A, B = gen_msa.split_family(ragged(fam, 5), range(0, 12), range(12, 17))
    galign_case("galign_gpf_lcl15", A, B, wt=1, lcl=15)
    galign_case("galign_gpf_lcl5_tgapf05", A, B, wt=1, lcl=5, tgapf=0.5)
    galign_case("galign_gpf_ragged_global", A, B, wt=1)
    A, B = gen_msa.split_family(ragged(fam, 6), range(0, 10), [10])
    galign_case("galign_hlf_lcl10", A, B, lcl=10)
    # NTV_ALB (DPunit_nv): small groups of raw residues, no profile (2 nj + ni < 8)
    A, B = gen_msa.split_family(hh, range(0, 2), range(2, 4))
    galign_case("galign_ntv_2x2", A, B)
    galign_case("galign_ntv_2x2_wt_f32", A, B, flavour="f", wt=1)
    A, B = gen_msa.split_family(hh, range(0, 3), [3])
    galign_case("galign_ntv_3x1_wt", A, B, wt=1)
    A, B = gen_msa.split_family(hh, [4], range(5, 8))
    galign_case("galign_ntv_1x3_twopiece", A, B, ls=3, mtx="blosum62")
    gl = gen_msa.synth_msa(7, 80, 0.2, 0.6, 43, gapless=True)
    galign_case("galign_ngp_gapless4x3", gl[:4], gl[4:], mtx="blosum62")          # NGP with thickness
    dn = gen_msa.synth_msa(12, 120, 0.05, 0.35, 44, dna=True)
    A, B = gen_msa.split_family(dn, range(0, 7), range(7, 12))
    galign_case("galign_dna_gpf_twopiece", A, B, molc="n", ls=3, wt=1)


def galign_rect_cases():
    """alignC over the whole rectangle (algmode.bnd = 0 -> alnmode NGP_ALN, forwardA + initA) for the groups that
    have no gap profile: gapless groups with thickness and single sequences."""
    import gen_msa
    gl = gen_msa.synth_msa(7, 80, 0.2, 0.6, 43, gapless=True)
    galign_case("galign_rect_ngp_gapless4x3", gl[:4], gl[4:], mtx="blosum62", bnd=0)
    galign_case("galign_rect_ngp_gapless4x3_twopiece", gl[:4], gl[4:], mtx="blosum62", ls=3, bnd=0)
    galign_case("galign_rect_ngp_gapless3x4_wt_f32", gl[4:], gl[:4], flavour="f", wt=1, bnd=0)
    p = gen_synth.synth_set(10, 120, 0.1, 0.7, 11)
    rag = [s[a:len(s) - b] for s, a, b in zip(gen_synth.synth_set(10, 160, 0.1, 0.6, 41),
                                             [0, 30, 0, 45, 10, 0, 60, 5, 0, 25], [0, 0, 40, 20, 0, 55, 0, 35, 15, 0])]
    galign_case("galign_rect_single_p01", [p[0]], [p[1]], flavour="f", mtx="blosum62", bnd=0)     # path opens on the diagonal
    galign_case("galign_rect_single_p24_twopiece", [p[2]], [p[4]], ls=3, bnd=0)                   # path opens with a gap
    galign_case("galign_rect_single_rag03_twopiece_u1", [rag[0]], [rag[3]], mtx="blosum62", ls=3, u1=1, bnd=0)
    galign_case("galign_rect_single_rag62", [rag[6]], [rag[2]], mtx="blosum62", bnd=0)            # 100 x 120, far off the main diagonal


def galign_swg_cases():
    """Smith-Waterman on groups (algmode.lcl = 16, mlt = 1): swg1st -> Fwd2c<SwgDPunit*>::forwardC gives the best local
    score and its box (colony 0); swg2nd aligns inside the box (align2).  The goldens also hold the banded global
    results of the same inputs."""
    import gen_msa
    fam = gen_msa.synth_msa(64, 90, 0.1, 0.5, 41)
    hh = gen_msa.synth_msa(40, 80, 0.3, 0.8, 45, indel_events=12.0)
    A, B = gen_msa.split_family(fam, range(0, 12), range(12, 17))
    galign_case("galign_swg_gpf_prof12_raw5_wt", A, B, wt=1, lcl=16)
    A, B = gen_msa.split_family(fam, range(0, 8), range(8, 14))
    galign_case("galign_swg_gpf_twopiece", A, B, ls=3, wt=1, lcl=16)
    A, B = gen_msa.split_family(fam, range(0, 10), [10])
    galign_case("galign_swg_hlf_prof10_single", A, B, lcl=16)
    galign_case("galign_swg_rhf_single_prof10_f32", B, A, flavour="f", wt=1, lcl=16)
    A, B = gen_msa.split_family(hh, range(0, 22), range(22, 40))
    galign_case("galign_swg_gpf_highhetero", A, B, wt=1, sh=-40, lcl=16)
    # (naive groups, NTV_ALB: algmode.lcl = 16 gives them nil ends, which the library's DPunit_nv form does not take)
    gl = gen_msa.synth_msa(7, 80, 0.2, 0.6, 43, gapless=True)
    galign_case("galign_swg_ngp_gapless4x3", gl[:4], gl[4:], mtx="blosum62", lcl=16)
    rnd = gen_synth.synth_set(6, 150, 0.9, 0.9, 77)                                # unrelated: a small island, many resets
    galign_case("galign_swg_single_unrelated", [rnd[0]], [rnd[4]], flavour="f", mtx="blosum62", lcl=16)
    rag = [s[a:len(s) - b] for s, a, b in zip(gen_synth.synth_set(10, 160, 0.1, 0.6, 41),
                                             [0, 30, 0, 45, 10, 0, 60, 5, 0, 25], [0, 0, 40, 20, 0, 55, 0, 35, 15, 0])]
    galign_case("galign_swg_single_rag03_twopiece", [rag[0]], [rag[3]], mtx="blosum62", ls=3, u1=1, sh=-20, lcl=16)


def alignb_case(name, seqs, flavour="f", **kv):
    """Aln2b1: alignB_ng (stdskl-normalised corner list) + HomScoreB_ng per pair."""
    os.makedirs(TMP, exist_ok=True)
    fa = os.path.join(TMP, name + ".fa")
    gen_synth.write_fasta(fa, seqs)
    al = refio.run("alignb", fa, flavour=flavour, **kv)
    mt = refio.run("matrix", fa, flavour=flavour, **kv)
    n = len(seqs)
    pairs = [dict(i=i, j=j, score=al["alignb"][(i, j)]["score"], hom=al["alignb"][(i, j)]["hom"],
                  skl=al["alignb"][(i, j)]["skl"]) for j in range(1, n) for i in range(j)]
    rec = dict(name=name, kind="alignb", flavour=flavour, params=al["header"], args=kv, seqs=seqs,
               matrix=mt["matrix"].tolist(), pairs=pairs)
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(rec, f)
    print("wrote", name, "pairs", len(pairs))


def alignb_cases():
    p12 = gen_synth.synth_set(12, 120, 0.1, 0.7, 11)
    alignb_case("alignb_p12_blosum62", p12)
    alignb_case("alignb_p12_pam_f64", p12, flavour="d", mtx="pam")
    alignb_case("alignb_p12_twopiece_f64", p12, flavour="d", ls=3)
    alignb_case("alignb_p12_sh3_u3v11", gen_synth.synth_set(12, 150, 0.2, 0.9, 12), sh=3, u=3, v=11)
    alignb_case("alignb_long700", gen_synth.synth_set(4, 700, 0.1, 0.5, 31))
    # terminal gaps: reduced penalty at true ends (initB_ng / lastB_ng with tgapf < 1), free ends per side
    rag = [s[a:len(s) - b] for s, a, b in zip(gen_synth.synth_set(10, 160, 0.1, 0.6, 41),
                                             [0, 30, 0, 45, 10, 0, 60, 5, 0, 25], [0, 0, 40, 20, 0, 55, 0, 35, 15, 0])]
    alignb_case("alignb_rag10_tgapf05_f64", rag, flavour="d", tgapf=0.5)
    alignb_case("alignb_rag10_tgapf0", rag, tgapf=0)
    alignb_case("alignb_rag10_lcl15_f64", rag, flavour="d", lcl=15)
    alignb_case("alignb_rag10_lcl6", rag, lcl=6)
    alignb_case("alignb_rag10_lcl9_twopiece_f64", rag, flavour="d", lcl=9, ls=3)


def sample_pair():
    """C1: the sample/pas ce13a1 x ce13a2 pair of sample/test.sh (annotation lines stripped)."""
    out = []
    for nm in ("ce13a1", "ce13a2"):
        path = os.path.join("/root/reference/sample/pas", nm)
        s = []
        with open(path) as f:
            for line in f:
                if line[0] in ">;#":
                    continue
                s.append("".join(c for c in line.strip() if c.isalpha()))
        out.append("".join(s))
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    p24 = gen_synth.synth_set(24, 120, 0.1, 0.6, 11)
    score_case("score_p24_blosum62", p24)
    score_case("score_p24_sh20", p24, sh=-20)
    score_case("score_p24_sh3", gen_synth.synth_set(24, 150, 0.2, 0.9, 12), sh=3)
    score_case("score_p24_sh0", gen_synth.synth_set(24, 150, 0.2, 0.9, 12), sh=0)
    score_case("score_p24_u3v11", p24, u=3, v=11)
    score_case("score_p24_tgapf05", p24, tgapf=0.5)
    score_case("score_p24_tgapf0", p24, tgapf=0.0, sh=-30)
    score_case("score_p24_pam_f32", p24, mtx="pam")
    score_case("score_p24_pam_f64", p24, flavour="d", mtx="pam")
    # ragged lengths, including very short sequences
    rag = [s[:k] for s, k in zip(gen_synth.synth_set(20, 300, 0.1, 0.7, 21),
                                 [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 300, 17, 64, 65, 31, 32, 33, 250])]
    score_case("score_ragged", rag)
    score_case("score_c2_first40", gen_synth.config_set("c2", 40))
    score_case("score_c5a_first40", gen_synth.config_set("c5a", 40))
    long_ = gen_synth.synth_set(6, 1300, 0.1, 0.5, 31)
    score_case("score_long1300", long_)
    score_case("score_c1_ce13a", sample_pair(), sh=-50)
    lcl_cases(p24, rag)
    # alignments with path (align2 -> alignC<DPunit> -> stdskl)
    p16 = p24[:16]
    align_case("align_p16_blosum62", p16)
    align_case("align_p16_sh3", gen_synth.synth_set(16, 150, 0.2, 0.9, 12), sh=3)
    align_case("align_p16_u3v11", p16, u=3, v=11)
    align_case("align_p16_twopiece_u1_1", p16, ls=3, u1=1)
    align_case("align_p16_twopiece_default", p16, ls=3)
    align_case("align_p16_twopiece_f64", p16, flavour="d", ls=3)
    align_case("align_p16_pam_f32", p16, mtx="pam")
    align_case("align_ragged", rag[:14])
    align_case("align_c2_first12", gen_synth.config_set("c2", 12))
    align_case("align_long1300", long_[:4])
    align_case("align_c1_ce13a", sample_pair(), sh=-50)
    galign_cases()
    galign_rect_cases()
    galign_swg_cases()
    alignb_cases()
    dna_pair_cases()


if __name__ == "__main__":
    if not refio.available("f"):
        sys.exit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    if len(sys.argv) > 1 and sys.argv[1] == "alignb":
        alignb_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "dna":
        dna_pair_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "galign":
        galign_cases()
        galign_rect_cases()
        galign_swg_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "rect":
        galign_rect_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "swg":
        galign_swg_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "lcl":      # only the lcl cases
        p24_ = gen_synth.synth_set(24, 120, 0.1, 0.6, 11)
        rag_ = [s[:k] for s, k in zip(gen_synth.synth_set(20, 300, 0.1, 0.7, 21),
                                      [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 300, 17, 64, 65, 31, 32, 33, 250])]
        lcl_cases(p24_, rag_)
    else:
        main()
