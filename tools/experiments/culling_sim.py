import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/cbf-inc-rrt_b200')
from oracle import oracle_f64 as O
from cbfrrt import workloads as W
models=O.load_models()
def ritter(P,r):
    # small enclosing ball of spheres (centres P, radii r): iterate
    c=P.mean(0)
    for _ in range(200):
        d=np.linalg.norm(P-c,axis=1)+r; i=d.argmax(); c=c+(P[i]-c)*0.02
    return c,(np.linalg.norm(P-c,axis=1)+r).max()
def sim(cfg, nstates=2048, thr=0.02, interp=False):
    w=W.config(cfg, scale=1e-9)
    robot=w["robot"]; m=models[robot]
    q=W.sample_states(robot,nstates,0)
    if interp:   # consecutive lanes = consecutive points of an edge
        E=nstates//32; a=W.sample_states(robot,E,0); b=a+np.random.default_rng(1).uniform(-0.35,0.35,a.shape)
        q=(a[:,None,:]+(np.arange(32)/32)[None,:,None]*(b-a)[:,None,:]).reshape(-1,a.shape[1])
    boxes=w["boxes"]; sph=m["spheres"]
    lk=np.array([s["link"] for s in sph]); rad=np.array([s["r"] for s in sph]); cl=np.array([s["c"] for s in sph])
    links=sorted(set(lk))
    Cs=np.zeros((len(q),len(sph),3)); Fr=[]
    for b in range(len(q)):
        fr=O.fk(m,q[b]); Cs[b]=O.sphere_centres(m,fr); Fr.append(fr)
    def sdb(P,c,h):
        a=np.abs(P-c)-h
        return np.linalg.norm(np.maximum(a,0),axis=-1)+np.minimum(a.max(-1),0)
    link0_floor={"panda":0.141,"magician":0.048}[robot]
    best=np.full(len(q),link0_floor)
    tot=0; done=0; lane=0
    for L in links:
        if L==-1: continue
        idx=np.where(lk==L)[0]
        cloc,R=ritter(cl[idx],rad[idx])
        cen=np.array([Fr[b][L][:3,:3]@cloc+Fr[b][L][:3,3] for b in range(len(q))])
        if L>0: best=np.minimum(best,(Cs[:,idx,2]-rad[idx]).min(1))   # floor pairs first
        T=np.maximum(best,thr)
        newbest=best.copy()
        for bx in boxes:
            lb=sdb(cen,bx[:3],bx[3:])-R
            need=lb<T
            nw=need.reshape(-1,32).any(1)
            tot+=len(idx)*len(nw); done+=len(idx)*nw.sum()+len(nw); lane+=len(idx)*need.mean()*len(nw)
            for i in idx:
                e=sdb(Cs[:,i],bx[:3],bx[3:])-rad[i]
                newbest=np.minimum(newbest,np.where(need,e,np.inf))
        best=newbest
    print(w["name"],"interp" if interp else "","R per link ok; warp-level executed:",round(done/tot,3),"lane-level:",round(lane/tot,3))
sim(1); sim(3); sim(0); sim(2); sim(2,interp=True); sim(4,nstates=1024)
