import sys, subprocess, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import c_oracle
def run(n,S,p,B=20000,seed=None,maxit=100,tol=1e-9):
    rng=np.random.default_rng(100+n*S if seed is None else seed)
    lo,hi=-np.ones(n,np.float32),np.ones(n,np.float32)
    a=rng.normal(size=(B,S,n)).astype(np.float32); c=rng.normal(scale=.8,size=(B,S)).astype(np.float32); ur=rng.uniform(-1.5,1.5,size=(B,n)).astype(np.float32)
    open('/tmp/qpm_in.bin','wb').write(a.tobytes()+c.tobytes()+ur.tobytes())
    subprocess.check_call(['tools/experiments/qp_multi_host',str(n),str(S),str(B),str(p),str(maxit),str(tol),'/tmp/qpm_in.bin','/tmp/qpm_out.bin'])
    raw=open('/tmp/qpm_out.bin','rb').read()
    u=np.frombuffer(raw[:B*n*4],np.float32).reshape(B,n); r=np.frombuffer(raw[B*n*4:B*n*4+B*S*4],np.float32).reshape(B,S); it=np.frombuffer(raw[B*n*4+B*S*4:],np.int32)
    uo,ro,ito=c_oracle.qp_multi(a,c,ur,lo,hi,p,want_iters=True)
    return a,c,ur,u,r,it,uo,ro,ito
if __name__=="__main__":
    for n,S,p in ((4,2,50.0),(7,3,1e3),(7,4,0.05),(7,2,1e9),(4,3,5000.),(7,1,50.),(4,4,1e9)):
        a,c,ur,u,r,it,uo,ro,ito=run(n,S,p, seed=9 if (n,S)==(4,4) else None)
        pc=min(p,1e6)
        err=np.abs(u-uo).max(1)
        obj=lambda uu,rr:((uu.astype(np.float64)-ur)**2).sum(1)+pc*rr.astype(np.float64).sum(1)
        do=obj(u,r)-obj(uo,ro)
        print(n,S,p,"iters pct",np.percentile(it,[50,99,100]),"nonconv",np.mean(it>=100),"err pct",np.percentile(err,[50,99,99.9,100]),"obj excess max",do.max(), "rel", (do/(1+obj(uo,ro))).max())
