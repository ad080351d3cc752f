import numpy as np, sys
sys.path.insert(0,'/root/repo/tools/scratch')
from qp_multi_proto import qp_mu
def solve(a,c,ur,lo,hi,p,tol,maxs=2000,dt=np.float64,eps=1e-6):
    a=a.astype(dt);c=c.astype(dt);ur=ur.astype(dt); lo=lo.astype(dt); hi=hi.astype(dt); p=dt(p)
    S,n=a.shape; mu=np.zeros(S,dt); t=ur.copy()
    for sw in range(maxs):
        moved=dt(0)
        for i in range(S):
            ti=t+dt(0.5)*mu[i]*a[i]
            m=dt(qp_mu(a[i],ti,lo,hi,c[i],p))
            d=abs(m-mu[i])*dt(0.5)*np.abs(a[i]).max()
            moved=max(moved,d); mu[i]=m; t=ti-dt(0.5)*m*a[i]
        if moved<=tol: return sw+1,mu,np.clip(t,lo,hi)
        if S>=2:
            u=np.clip(t,lo,hi); g=c+a@u
            W=((mu>0)&(mu<p))|((mu<=0)&(g>0))|((mu>=p)&(g<0)); F=(t>lo)&(t<hi)
            if W.sum()>=2:
                AW=a[W][:,F]
                M=dt(0.5)*AW@AW.T
                M=M+np.eye(W.sum(),dtype=dt)*(dt(eps)*np.trace(M)+dt(1e-30))
                dW=np.linalg.solve(M,g[W]).astype(dt)
                d=np.zeros(S,dt); d[W]=dW
                # drop components pointing out of the box
                d[(mu<=0)&(d<0)]=0; d[(mu>=p)&(d>0)]=0
                if not np.any(d!=0): continue
                with np.errstate(divide='ignore',invalid='ignore'):
                    amax=np.min(np.where(d>0,(p-mu)/d,np.where(d<0,-mu/d,np.inf)))
                w=a.T@d
                al=dt(qp_mu(w,t,lo,hi,d@c,amax))
                # qp_mu's constant: g(alpha)= cc + w.clip(t-0.5 alpha w)
                mu=np.clip(mu+al*d,0,p); t=t-dt(0.5)*al*w
                moved=max(moved,al*dt(0.5)*np.abs(w).max())
    return maxs,mu,np.clip(t,lo,hi)
if __name__=="__main__":
    rng=np.random.default_rng(108)
    for dt in (np.float64,np.float32):
      for n,S,p in ((4,2,50.0),(7,3,1e3),(7,4,0.05),(4,4,1e6),(7,2,1e6)):
        B=2000
        lo,hi=-np.ones(n),np.ones(n)
        a=rng.normal(size=(B,S,n)); c=rng.normal(scale=.8,size=(B,S)); ur=rng.uniform(-1.5,1.5,size=(B,n))
        sw=np.array([solve(a[b],c[b],ur[b],lo,hi,p,1e-6,dt=dt)[0] for b in range(B)])
        print(dt.__name__,n,S,p,np.percentile(sw,[50,90,99,99.9,100]))
def debug():
    rng=np.random.default_rng(5)
    n,S,p=4,4,1e6
    B=3000
    lo,hi=-np.ones(n),np.ones(n)
    a=rng.normal(size=(B,S,n)); c=rng.normal(scale=.8,size=(B,S)); ur=rng.uniform(-1.5,1.5,size=(B,n))
    bad=[b for b in range(B) if solve(a[b],c[b],ur[b],lo,hi,p,1e-6,maxs=100)[0]==100]
    print(len(bad))
    b=bad[0]
    for ms in (5,20,21,22,100):
        sw,mu,u=solve(a[b],c[b],ur[b],lo,hi,p,1e-6,maxs=ms)
        print(ms,mu,u,c[b]+a[b]@u)
