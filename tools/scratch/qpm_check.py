import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import c_oracle
from conftest import qp_kkt_residual
rng=np.random.default_rng(108)
for n,S,p in ((4,2,50.0),(7,3,1e3),(7,4,0.05),(4,4,1e9),(7,2,1e9),(4,3,5000.)):
    B=4000
    lo,hi=-np.ones(n,np.float32),np.ones(n,np.float32)
    a=rng.normal(size=(B,S,n)).astype(np.float32); c=rng.normal(scale=.8,size=(B,S)).astype(np.float32); ur=rng.uniform(-1.5,1.5,size=(B,n)).astype(np.float32)
    u,r,it=c_oracle.qp_multi(a,c,ur,lo,hi,p,want_iters=True)
    conv=it<2000
    res=[qp_kkt_residual(a[b],c[b],ur[b],lo,hi,min(p,1e6),u[b]) for b in range(0,B,10) if conv[b]]
    print(n,S,p,"conv",conv.mean(),"iters pct",np.percentile(it,[50,99,100]),"kkt max",max(res))
