"""Run oracle/_ref/ref_driver_{f,d} (the unmodified reference) and parse its text output.
TEST INFRASTRUCTURE: used by tools/make_golden.py, tests/ (when oracle/_ref exists) and bench.py's
cpu_baseline / --impl reference legs."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def driver(flavour="f", gpu=False):
    """gpu=True: the same driver with shim/shim_fwd2d1.cc linked in front of the reference's fwd2d1.o, i.e.
    the unmodified reference calling libprrn_gpu.so for alnScoreD (oracle/Makefile: ref_gpu)."""
    return os.path.join(REFDIR, "ref_driver_" + flavour + ("_gpu" if gpu else ""))


def available(flavour="f"):
    return os.path.exists(driver(flavour)) and os.path.exists(os.path.join(REFDIR, "table", "blosum62"))


def run(cmd, fasta, flavour="f", timeout=3600, gpu=False, **kv):
    env = dict(os.environ, ALN_TAB=os.path.join(REFDIR, "table"))
    args = [driver(flavour, gpu), cmd, fasta] + ["%s=%s" % (k, v) for k, v in kv.items()]
    out = subprocess.run(args, env=env, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("ref_driver failed (%d): %s" % (out.returncode, out.stderr[-2000:]))
    return parse(out.stdout)


def parse(text):
    r = {"header": None, "scores": {}, "ends": {}, "dist": [], "aligns": {}, "fstat": {}, "time": None,
         "matrix": None, "seqs": {}}
    lines = text.splitlines()
    i = 0
    while i < len(lines):
        t = lines[i].split()
        i += 1
        if not t:
            continue
        if t[0].startswith("#ref_driver"):
            r["header"] = dict(x.split("=", 1) for x in t[1:])
        elif t[0] == "score":
            r["scores"][(int(t[1]), int(t[2]))] = float(t[3])
            if len(t) > 4:
                r["ends"][(int(t[1]), int(t[2]))] = (int(t[4]), int(t[5]))
        elif t[0] == "dist":
            r["dist"].append(float(t[2]))
        elif t[0] == "time":
            r["time"] = float(t[1])
        elif t[0] == "dim":
            dim = int(t[1])
            r["matrix"] = np.array([[float(x) for x in lines[i + k].split()] for k in range(dim)])
            i += dim
        elif t[0] == "seq":
            r["seqs"][int(t[1])] = dict(len=int(t[2]), left=int(t[3]), right=int(t[4]),
                                         codes=[int(x) for x in t[6:]])
        elif t[0] == "align":
            key = (int(t[1]), int(t[2]))
            swp = int(t[3].split("=")[1])
            mode = int(t[4].split("=")[1])
            scr = float(t[5])
            skl = None
            if t[6] == "skl" and t[7] != "0":
                n = int(t[7])
                flag = int(t[8])
                v = [int(x) for x in t[10:10 + 2 * n]]
                skl = dict(n=n, flag=flag, pts=list(zip(v[0::2], v[1::2])))
            r["aligns"][key] = dict(swp=swp, mode=mode, score=scr, skl=skl)
        elif t[0] == "alignb":
            key = (int(t[1]), int(t[2]))
            skl = None
            if t[5] == "skl" and t[6] != "0":
                n = int(t[6])
                v = [int(x) for x in t[9:9 + 2 * n]]
                skl = list(zip(v[0::2], v[1::2]))
            r.setdefault("alignb", {})[key] = dict(score=float(t[3]), hom=float(t[4]), skl=skl)
        elif t[0] == "fstat":
            r["fstat"][(int(t[1]), int(t[2]))] = [float(x) for x in t[3:]]
    r["dist"] = np.array(r["dist"])
    return r


def parse_galign(text):
    """Parse `ref_driver galign` output: the staged inputs Fwd2c reads per column + the alignC result."""
    r = {"groups": [None, None]}
    lines = text.splitlines()
    i = 0
    cur = None
    while i < len(lines):
        t = lines[i].split()
        i += 1
        if not t:
            continue
        if t[0].startswith("#ref_driver"):
            r["header"] = dict(x.split("=", 1) for x in t[1:])
        elif t[0] == "pwdm":
            r["pwdm"] = {t[k]: int(t[k + 1]) for k in range(1, len(t), 2)}
        elif t[0] == "pwdc":
            r["pwdc"] = {t[k]: float(t[k + 1]) for k in range(1, len(t), 2)}
        elif t[0] == "dim":
            dim = int(t[1])
            r["matrix"] = [[float(x) for x in lines[i + k].split()] for k in range(dim)]
            i += dim
        elif t[0] == "group":
            cur = {t[k]: (float(t[k + 1]) if t[k] == "sumwt" else int(t[k + 1])) for k in range(2, len(t), 2)}
            cur.update(pos=[], cfq=[], dfq=[], efq=[], res=[], vss=[], sfq=[], tfq=[], rfq=[], weight=None)
            r["groups"][int(t[1])] = cur
        elif t[0] == "weight":
            n = int(t[1])
            cur["weight"] = [float(x) for x in t[2:2 + n]] if n else None
        elif t[0] == "pos":
            cur["pos"].append(int(t[1]))
            k = 3
            cur["cfq"].append(float(t[k])); cur["dfq"].append(float(t[k + 1])); cur["efq"].append(float(t[k + 2]))
            k += 3
            assert t[k] == "res"
            cur["res"].append([int(x) for x in t[k + 1:k + 1 + cur["many"]]])
            k += 1 + cur["many"]
            assert t[k] == "vss"
            nv = int(t[k + 1])
            cur["vss"].append([float(x) for x in t[k + 2:k + 2 + nv]])
            k += 2 + nv
            for tag in ("sfq", "tfq", "rfq"):
                assert t[k] == tag, (t[k], tag)
                n = int(t[k + 1])
                k += 2
                if n < 0:
                    cur[tag].append(None)
                else:
                    cur[tag].append([[int(t[k + 3 * q]), float(t[k + 3 * q + 1]), int(t[k + 3 * q + 2])] for q in range(n)])
                    k += 3 * n
        elif t[0] == "window":
            r["window"] = [int(x) for x in t[1:4]]
        elif t[0] == "time":
            r["time"] = float(t[1])
        elif t[0] == "mt":
            pts = None
            if t[3] == "skl" and t[4] != "0":
                n = int(t[4])
                v = [int(x) for x in t[7:7 + 2 * n]]
                pts = [list(x) for x in zip(v[0::2], v[1::2])]
            r.setdefault("mt", []).append(dict(score=float(t[2]), skl=pts))
        elif t[0] == "homscore":
            r["homscore"] = dict(score=float(t[1]), rr=[int(t[2]), int(t[3])])
        elif t[0] == "swg":
            if t[1] == "none":
                r["swg"] = None
            else:
                r["swg"] = dict(size=int(t[2]), val=float(t[4]), **{t[k]: int(t[k + 1]) for k in range(5, len(t), 2)})
        elif t[0] == "swg2nd":
            pts = None
            if t[2] == "skl" and t[3] != "0":
                n = int(t[3])
                v = [int(x) for x in t[6:6 + 2 * n]]
                pts = [list(x) for x in zip(v[0::2], v[1::2])]
            r["swg2nd"] = dict(score=float(t[1]), skl=pts)
        elif t[0] in ("alignc", "align2"):
            off = 1 if t[0] == "alignc" else 3
            scr = float(t[off])
            pts = None
            if t[off + 1] == "skl" and t[off + 2] != "0":
                n = int(t[off + 2])
                v = [int(x) for x in t[off + 5:off + 5 + 2 * n]]
                pts = [list(x) for x in zip(v[0::2], v[1::2])]
            r[t[0]] = dict(score=scr, skl=pts)
            if t[0] == "align2":
                r["align2"]["sh"] = int(t[2])
    return r


def run_galign(fa, fb, flavour="d", timeout=3600, **kv):
    env = dict(os.environ, ALN_TAB=os.path.join(REFDIR, "table"))
    args = [driver(flavour), "galign", fa, "fb=" + fb] + ["%s=%s" % (k, v) for k, v in kv.items()]
    out = subprocess.run(args, env=env, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("ref_driver galign failed (%d): %s" % (out.returncode, out.stderr[-2000:]))
    return parse_galign(out.stdout)
