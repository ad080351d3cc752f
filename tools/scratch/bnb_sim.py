import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/cbf-inc-rrt_b200')
from oracle import oracle_f64 as O
from cbfrrt import workloads as W
models=O.load_models()
def sim(cfg, nstates=1024, per_sphere=False):
    w=W.config(cfg, scale=1e-9)
    robot=w["robot"]; m=models[robot]
    q=W.sample_states(robot,nstates,0)
    boxes=w["boxes"]; 
    sph=m["spheres"]; links=sorted(set(s["link"] for s in sph))
    Cs=np.zeros((nstates,len(sph),3))
    for b in range(nstates):
        Cs[b]=O.sphere_centres(m,O.fk(m,q[b]))
    rad=np.array([s["r"] for s in sph]); lk=np.array([s["link"] for s in sph])
    def sdb(P,c,h):
        a=np.abs(P-c)-h
        return np.linalg.norm(np.maximum(a,0),axis=-1)+np.minimum(a.max(-1),0)
    best=np.full(nstates,np.inf)
    if w["floor"]:
        best=np.minimum(best,(Cs[:,:,2]-rad)[:,lk>0].min(1))   # rough
    tot=0; done=0; done_lane=0
    for L in links:
        if L==links[0]: continue   # base link static
        idx=np.where(lk==L)[0]
        cen=Cs[:,idx].mean(1); R=(np.linalg.norm(Cs[:,idx]-cen[:,None],axis=-1)+rad[idx]).max(1)
        newbest=best.copy()
        for bx in boxes:
            lb=sdb(cen,bx[:3],bx[3:])-R
            need=lb<best
            nw=need.reshape(-1,32).any(1)
            tot+=len(idx)*len(nw); done+=len(idx)*nw.sum()+ len(nw)*1.0  # +1 pair-equivalent for the bound test
            done_lane+=len(idx)*need.mean()*len(nw)
            for i in idx:
                e=sdb(Cs[:,i],bx[:3],bx[3:])-rad[i]
                newbest=np.minimum(newbest,np.where(need,e,np.inf))
        best=newbest
    print(w["name"],"pairs fraction executed (warp-level incl. test):",done/tot,"lane-level:",done_lane/tot)
sim(1); sim(3); sim(0); sim(2)
