"""Where the host-buffer pg_calcdist call spends its time on C2 (PG_TIMING=1 prints the library's own breakdown)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import torch
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode
import gen_synth
seqs = gen_synth.synth_set(1000, 400, 0.1, 0.6, 1)
enc = [seqcode.encode_protein(s) for s in seqs]
ss = P.SeqSet(enc)
M = np.array(json.load(open(os.path.join(ROOT, "tests/golden/score_p24_blosum62.json")))["matrix"])
prm = P.Params(P.ALPRM(sh=-60), vtype=1)
ctx = P.Context(0)
npair = ss.n * (ss.n - 1) // 2
out = torch.empty(npair, dtype=torch.float64).pin_memory().numpy()
for _ in range(3):
    ctx.calcdist(ss, prm, M, 0, npair, out=out)
ts = []
for _ in range(10):
    t = time.perf_counter(); ctx.calcdist(ss, prm, M, 0, npair, out=out); ts.append(time.perf_counter() - t)
print("host-buffer call: median %.3f ms, min %.3f ms" % (1e3 * np.median(ts), 1e3 * min(ts)))
d = ctx.upload(ss)
dout = torch.empty(npair, dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.calcdist_dev(d, prm, M, 0, npair, dout.data_ptr()); torch.cuda.synchronize()
ts = []
for _ in range(10):
    t = time.perf_counter(); ctx.calcdist_dev(d, prm, M, 0, npair, dout.data_ptr()); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
print("device-resident call (+sync): median %.3f ms" % (1e3 * np.median(ts)))
