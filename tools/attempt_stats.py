import sys; sys.path.insert(0,'max-decoy_b200'); sys.path.insert(0,'tests')
import numpy as np, maxdecoy
from maxdecoy import synth, SearchParams
prots=synth.synthetic_proteins(20000)
sp,_=synth.synthetic_spectra(prots,2000,2,seed=7)
e=maxdecoy.Engine(); e.digest(prots,2,5,50); e.set_modifications([synth.CAM],0); e.index_build()
psms,st=e.identify(sp,SearchParams(10,10,n_decoys=1000,seed=20260101,keep_decoys=True))
d=e.last_decoys(); off=d['off'].astype(np.int64); att=d['attempt'].astype(np.int64)
need=np.array([att[off[s+1]-1]+1 if off[s+1]>off[s] else 0 for s in range(len(sp))])
cnt=np.diff(off)
print('attempts used per spectrum', st['n_attempts']/len(sp), 'needed (index of last accepted +1) mean', need.mean(), 'p50', np.median(need), 'p90', np.percentile(need,90), 'max', need.max())
print('full spectra', (cnt==1000).mean(), 'yield', (cnt/np.maximum(need,1)).mean())
M=sp.precursor_mz*sp.charge
for lo,hi in ((0,1200),(1200,2000),(2000,3000),(3000,6000)):
    m=(M>=lo)&(M<hi); print(lo,hi,'n',m.sum(),'need',need[m].mean())
