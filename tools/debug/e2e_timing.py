import os, sys, json, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import numpy as np
import gen_synth
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode
seqs = gen_synth.synth_set(1000, 400, 0.1, 0.6, 1)
enc = [seqcode.encode_protein(s) for s in seqs]
M = np.array(json.load(open('/root/repo/tests/golden/score_p24_blosum62.json'))['matrix'])
ctx = P.Context(0)
ss = P.SeqSet(enc); prm = P.Params()
for i in range(6):
    t0 = time.perf_counter(); d = ctx.calcdist(ss, prm, M); print('call %.3f ms kernel %.3f' % (1e3 * (time.perf_counter() - t0), ctx.last_kernel_ms()), file=sys.stderr)
