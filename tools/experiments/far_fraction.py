"""fraction of link-sphere positions whose grid cell is farther than R_MAX + T_MAX from every obstacle (mask-only skip)"""
import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/cbf-inc-rrt_b200')
from oracle import oracle_f64 as O
from cbfrrt import workloads as W
models=O.load_models()
def sdb(P,c,h):
    a=np.abs(P-c)-h
    return np.linalg.norm(np.maximum(a,0),axis=-1)+np.minimum(a.max(-1),0)
for cfg in (2,4):
    w=W.config(cfg, scale=1e-9); robot=w["robot"]; m=models[robot]
    q=W.sample_states(robot,1500,5)
    C=np.concatenate([O.sphere_centres(m,O.fk(m,q[b])) for b in range(len(q))])
    rad=np.tile([s["r"] for s in m["spheres"]],len(q)); lk=np.tile([s["link"] for s in m["spheres"]],len(q))
    keep=lk>=0; C=C[keep]; rad=rad[keep]
    d=np.full(len(C),np.inf)
    for bx in w["boxes"]: d=np.minimum(d,sdb(C,bx[:3],bx[3:]))
    for sp in w["spheres"]: d=np.minimum(d,np.linalg.norm(C-sp[:3],axis=1)-sp[3])
    rmax=max(s["r"] for s in m["spheres"]); diag=0.05*np.sqrt(3)
    for tmax in (0.06,):
        far=d>rmax+tmax+diag
        print(w["name"],"rmax",round(rmax,3),"far fraction of sphere look-ups:",round(far.mean(),3), " per-state all-far:", round(far.reshape(len(q),-1).all(1).mean(),3))
