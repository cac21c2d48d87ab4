#!/usr/bin/env python
"""End-to-end `prrn` MSA runs (the "prrn MSA wall-s" part of BASELINE.json's metric): the reference's own prrn5
program, unmodified, (a) as it is -- every DP on the host CPU (oracle/_ref/prrn5_cpu) -- and (b) linked with shim/*.cc
so that alnScoreD, calcdist, alignC / HomScoreC and alignB_ng run in libprrn_gpu.so (oracle/_ref/prrn5_gpu).

    python tools/run_prrn.py [--arm cpu|gpu|both] [--freeze] CONFIG ...

CONFIG:
    small      40 x ~200 aa                      prrn5 -m blosum62
    mid        80 x ~300 aa                      prrn5 -m blosum62
    c3         200 x ~500 aa (BASELINE config 3) prrn5 -m blosum62
    c3t4       the MSA of c3 as a pre-aligned start, refined with B = 4 candidate partitions per cycle
               (Prrn::best_of_n, src/prrn5.cc:594): prrn5 -m blosum62 -t4 -r4   (needs c3's CPU output first)
    c4n50 / c4n60 / c4n100   the first 50 / 60 / 100 of the 100 DNA sequences of ~2 kb of BASELINE config 4, two-piece
               gap penalties: prrn5 -yl3  (the stock reference crashes above 60 DNA members, SURVEY.md N4: n100 is
               GPU-timed only)
Prints one JSON line per config: wall seconds per arm, md5 of the MSA body, and whether the arms agree with each other
and with the frozen result of the CPU arm (tests/golden/prrn_msa.json; --freeze writes it, in the container where
/root/reference was built)."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
FROZEN = os.path.join(ROOT, "tests", "golden", "prrn_msa.json")
BIN = {"cpu": "prrn5_cpu", "gpu": "prrn5_gpu"}

CONFIGS = {
    "small": dict(seqs=lambda: gen_synth.synth_set(40, 200, 0.1, 0.6, 3), args=["-m", "blosum62"]),
    "mid": dict(seqs=lambda: gen_synth.synth_set(80, 300, 0.1, 0.6, 2), args=["-m", "blosum62"]),
    "c3": dict(seqs=lambda: gen_synth.config_set("c3"), args=["-m", "blosum62"]),
    "c3t4": dict(start="c3", args=["-m", "blosum62", "-t4", "-r4"]),
    "c4n20": dict(seqs=lambda: gen_synth.config_set("c4", 20), args=["-yl3"]),
    "c4n50": dict(seqs=lambda: gen_synth.config_set("c4", 50), args=["-yl3"]),
    "c4n60": dict(seqs=lambda: gen_synth.config_set("c4", 60), args=["-yl3"]),
    "c4n100": dict(seqs=lambda: gen_synth.config_set("c4", 100), args=["-yl3"], gpu_only=True),
}


def msa_body(stdout):
    return "\n".join(l for l in stdout.splitlines() if not l.startswith(">") and "sec" not in l)


def run(arm, path, args, env_extra=None, timeout=3000):
    env = dict(os.environ, ALN_TAB=os.path.join(REFDIR, "table"))
    env.update(env_extra or {})
    t0 = time.perf_counter()
    out = subprocess.run([os.path.join(REFDIR, BIN[arm])] + args + [path], env=env, capture_output=True, text=True,
                         timeout=timeout)
    dt = time.perf_counter() - t0
    r = {"wall_s": dt, "rc": out.returncode, "msa_md5": hashlib.md5(msa_body(out.stdout).encode()).hexdigest(),
         "lines": len(out.stdout.splitlines())}
    if out.returncode:
        r["stderr"] = out.stderr[-400:]
    stats = [l for l in out.stderr.splitlines() if l.startswith("prrn_gpu")]
    if stats:
        r["stats"] = stats
    return r, out.stdout


def input_path(name):
    c = CONFIGS[name]
    path = "/tmp/prrn_in_%s.fa" % name
    if "start" in c:        # pre-aligned start = the frozen MSA of another config (the CPU arm's own output)
        src = os.path.join(ROOT, "tests", "golden", "prrn_%s_cpu.msa" % c["start"])
        if not os.path.exists(src):
            return None
        return src
    gen_synth.write_fasta(path, c["seqs"]())
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["cpu", "gpu", "both"])
    ap.add_argument("--freeze", action="store_true", help="record the CPU arm's md5 (and keep c3's MSA as the start of c3t4)")
    ap.add_argument("configs", nargs="+")
    a = ap.parse_args()
    frozen = json.load(open(FROZEN)) if os.path.exists(FROZEN) else {}
    for name in a.configs:
        c = CONFIGS[name]
        path = input_path(name)
        res = {"config": name, "command": "prrn5 " + " ".join(c["args"]), "host_cores": os.cpu_count()}
        if path is None:
            res["unavailable"] = "pre-aligned start missing: run --freeze %s first" % c["start"]
            print(json.dumps(res), flush=True)
            continue
        arms = ["cpu", "gpu"] if a.arm == "both" else [a.arm]
        if c.get("gpu_only") and "cpu" in arms:
            arms.remove("cpu")
        for arm in arms:
            if not os.path.exists(os.path.join(REFDIR, BIN[arm])):
                res[arm] = "not built"
                continue
            env = {"PRRN_GPU_STATS": "1"} if arm == "gpu" else {}
            r, stdout = run(arm, path, c["args"], env)
            res[arm] = r
            if arm == "cpu" and a.freeze and r["rc"] == 0:
                frozen[name] = {"msa_md5": r["msa_md5"], "lines": r["lines"], "cpu_wall_s_here": round(r["wall_s"], 2),
                                "command": res["command"]}
                if name == "c3":
                    with open(os.path.join(ROOT, "tests", "golden", "prrn_c3_cpu.msa"), "w") as f:
                        f.write(stdout)
        if isinstance(res.get("cpu"), dict) and isinstance(res.get("gpu"), dict):
            res["identical_msa"] = res["cpu"]["msa_md5"] == res["gpu"]["msa_md5"]
        if name in frozen and isinstance(res.get("gpu"), dict):
            res["gpu_equals_frozen_cpu_msa"] = res["gpu"]["msa_md5"] == frozen[name]["msa_md5"]
            res["frozen_cpu_wall_s_container"] = frozen[name]["cpu_wall_s_here"]
        print(json.dumps(res), flush=True)
    if a.freeze:
        with open(FROZEN, "w") as f:
            json.dump(frozen, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
