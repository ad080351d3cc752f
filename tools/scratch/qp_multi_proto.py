import numpy as np
def qp_g(a,t,lo,hi,c,mu): return c + a @ np.clip(t-0.5*mu*a,lo,hi)
def qp_mu(a,t,lo,hi,c,p):
    g0=qp_g(a,t,lo,hi,c,0.0)
    if not g0>0: return 0.0
    bp=[0.0]
    for i in range(len(a)):
        if a[i]!=0:
            for m in (2*(t[i]-lo[i])/a[i], 2*(t[i]-hi[i])/a[i]):
                if 0<m<p: bp.append(m)
    bp.append(p); bp.sort()
    ml,gl=0.0,g0
    for m in bp[1:]:
        gh=qp_g(a,t,lo,hi,c,m)
        if gh<=0: return ml+(m-ml)*gl/(gl-gh) if gl>gh else m
        ml,gl=m,gh
    return p
def solve(a,c,ur,lo,hi,p,tol,newton=True,maxs=2000,dt=np.float64):
    a=a.astype(dt);c=c.astype(dt);ur=ur.astype(dt)
    S,n=a.shape; mu=np.zeros(S,dt); t=ur.copy()
    nacc=0
    for sw in range(maxs):
        moved=0
        for i in range(S):
            ti=t+dt(0.5)*mu[i]*a[i]
            m=dt(qp_mu(a[i],ti,lo,hi,c[i],p))
            d=abs(m-mu[i])*0.5*np.abs(a[i]).max()
            moved=max(moved,d); mu[i]=m; t=ti-dt(0.5)*m*a[i]
        if moved<=tol: return sw+1,mu,np.clip(t,lo,hi),nacc
        if newton:
            W=(mu>0)&(mu<p); F=(t>lo)&(t<hi)
            if W.sum()>=2:
                AW=a[W][:,F]
                M=dt(0.5)*AW@AW.T
                g=c[W]+a[W]@np.clip(t,lo,hi)
                M=M+np.eye(W.sum(),dtype=dt)*dt(1e-7)*np.trace(M)
                try:
                    d=np.linalg.solve(M,g).astype(dt)
                except np.linalg.LinAlgError:
                    continue
                mu2=mu.copy(); mu2[W]+=d
                t2=t-dt(0.5)*(a[W].T@d)
                ok=np.all((mu2[W]>=0)&(mu2[W]<=p)) and np.array_equal((t2>lo)&(t2<hi),F) and np.array_equal(t2>=hi,t>=hi)
                if ok:
                    mu,t=mu2,t2; nacc+=1
    return maxs,mu,np.clip(t,lo,hi),nacc
if __name__=="__main__":
    rng=np.random.default_rng(108)
    for n,S,p in ((4,2,50.0),(7,3,1e3),(7,4,0.05),(4,4,1e6)):
        B=2000
        lo,hi=-np.ones(n),np.ones(n)
        a=rng.normal(size=(B,S,n)); c=rng.normal(scale=.8,size=(B,S)); ur=rng.uniform(-1.5,1.5,size=(B,n))
        for newton in (False,True):
            sw=np.array([solve(a[b],c[b],ur[b],lo,hi,p,1e-6,newton)[0] for b in range(B)])
            print(n,S,p,newton,np.percentile(sw,[50,90,99,99.9,100]))
