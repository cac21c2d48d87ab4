import time, sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tools')
t0=time.perf_counter()
import numpy as np
import prrn_aln_b200 as P
from prrn_aln_b200 import seqcode
import gen_synth
t1=time.perf_counter()
L=P.load_library()
t2=time.perf_counter()
ctx=P.Context(0)
t3=time.perf_counter()
seqs=gen_synth.synth_set(2, 200, 0.1, 0.6, 3)
enc=[seqcode.encode_protein(s) for s in seqs]
M=np.array(json.load(open('/root/repo/tests/golden/score_p24_blosum62.json'))['matrix'])
ss=P.SeqSet(enc); prm=P.Params()
ts=[]
for i in range(5):
    a=time.perf_counter(); P.alnScoreD(ss, M, prm, pairs=[(0,1)], ctx=ctx) if False else ctx.calcdist(ss, prm, M); ts.append(time.perf_counter()-a)
print('import %.3f load_lib %.3f create %.3f calls %s' % (t1-t0, t2-t1, t3-t2, ['%.4f'%x for x in ts]))
