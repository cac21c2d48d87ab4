#!/usr/bin/env python
"""One shimmed prrn5 run with PRRN_GPU_STATS=1: wall seconds and where the alignC calls went."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_synth
n, length, seed = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "200x500x1").split("x"))
fa = "/tmp/prrn_stats_in.fa"
gen_synth.write_fasta(fa, gen_synth.synth_set(n, length, 0.1, 0.6, seed))
env = dict(os.environ, ALN_TAB=os.path.join(ROOT, "oracle", "_ref", "table"), PRRN_GPU_STATS="1")
t = time.perf_counter()
r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "prrn5_gpu"), "-m", "blosum62", fa], env=env, capture_output=True, text=True)
print("wall %.2f s rc %d" % (time.perf_counter() - t, r.returncode))
print([l for l in r.stderr.splitlines() if l.startswith("prrn_gpu")])
import hashlib
body = "\n".join(l for l in r.stdout.splitlines() if not l.startswith(">") and "sec" not in l)
print("msa md5", hashlib.md5(body.encode()).hexdigest())
