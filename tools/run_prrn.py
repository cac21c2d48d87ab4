#!/usr/bin/env python
"""End-to-end `prrn` MSA runs (the "prrn MSA wall-s" part of BASELINE.json's metric): the reference's own
prrn5 program, unmodified, (a) as it is -- every DP on the host CPU -- and (b) linked with shim/*.cc so
that alnScoreD and alignC<DPunit | DPunit_hf | DPunit_pf> run in libprrn_gpu.so (oracle/Makefile: prrn).
Prints one JSON line per input: wall seconds of both, and whether the two MSAs are identical.

Note what (b) is and is not: the shims are PER-CALL bindings.  prrn5's refinement is a chain of dependent
align2 calls, so each group alignment is a batch of ONE on the GPU (K3's latency mode); the batch entry
points (pg_calcdist, pg_align_groups over best_of_n candidates) need the two small source patches of
INTEGRATION.md and are not exercised here."""
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
EXTRA = os.environ.get("PRRN_ARGS", "").split()     # e.g. PRRN_ARGS="-t16": pthread workers (best_of_n, CalcServer)


def run(binary, fa, extra):
    env = dict(os.environ, ALN_TAB=os.path.join(REFDIR, "table"))
    t0 = time.perf_counter()
    out = subprocess.run([os.path.join(REFDIR, binary)] + extra + [fa], env=env, capture_output=True, text=True, timeout=3000)
    dt = time.perf_counter() - t0
    body = "\n".join(l for l in out.stdout.splitlines() if not l.startswith(">") and "sec" not in l)
    return dt, out.returncode, hashlib.md5(body.encode()).hexdigest(), out.stdout, out.stderr[-300:]


def main():
    cases = [(40, 200, 3), (200, 500, 1)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    for n, length, seed in cases:
        fa = "/tmp/prrn_in_%d_%d.fa" % (n, length)
        gen_synth.write_fasta(fa, gen_synth.synth_set(n, length, 0.1, 0.6, seed))
        res = {"config": "prrn5 %s %d x ~%d aa (seed %d)" % (" ".join(EXTRA), n, length, seed), "host_cores": os.cpu_count()}
        outs = {}
        for tag, binary in (("cpu", "prrn5_cpu"), ("gpu_shims", "prrn5_gpu")):
            if not os.path.exists(os.path.join(REFDIR, binary)):
                res[tag] = "not built"
                continue
            dt, rc, md5, out, err = run(binary, fa, ["-m", "blosum62"] + EXTRA)
            res[tag] = {"wall_s": dt, "rc": rc, "msa_md5": md5, "lines": len(out.splitlines())}
            if rc:
                res[tag]["stderr"] = err
            outs[tag] = out
        if len(outs) == 2:
            res["identical_msa"] = res["cpu"]["msa_md5"] == res["gpu_shims"]["msa_md5"]
        print(json.dumps(res))


if __name__ == "__main__":
    main()
