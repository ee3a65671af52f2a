"""GPU essential-matrix RANSAC against the ground-truth model on the synthetic two-view scenes of tests/test_gpu_geometry.py: MSAC cost
and inlier recall of xfeat_ransac_essential next to the true E, then a numpy replay of the local optimisation from the returned model.
    python tools/essential_check.py      (one GPU)"""
import numpy as np, cv2, sys, torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from tests.test_gpu_geometry import synth_two_view
from accelerated_features_b200.geometry import find_essential_batch
def proj_E(E):
    U,S,Vt=np.linalg.svd(E); return U@np.diag([1,1,0])@Vt
def rows(x,u): return np.stack([u[:,0]*x[:,0],u[:,0]*x[:,1],u[:,0],u[:,1]*x[:,0],u[:,1]*x[:,1],u[:,1],x[:,0],x[:,1],np.ones(len(x))],1)
def samp(E,x,u):
    xh=np.c_[x,np.ones(len(x))]; uh=np.c_[u,np.ones(len(u))]
    Ex=xh@E.T; Etu=uh@E; e=(uh*Ex).sum(1); den=Ex[:,0]**2+Ex[:,1]**2+Etu[:,0]**2+Etu[:,1]**2
    return e*e/den, den
def fit(x,u,w):
    A=rows(x,u)*w[:,None]; _,_,Vt=np.linalg.svd(A.T@A); return proj_E(Vt[-1].reshape(3,3))
rng=np.random.default_rng(3)
for n,frac in [(2000,.6),(1200,.45),(600,.7),(300,.5)]:
    p0,p1,inl,K,T=synth_two_view(rng,n,frac,0.5)
    Ki=np.linalg.inv(K); x=(np.c_[p0,np.ones(n)]@Ki.T)[:,:2].astype(np.float32); u=(np.c_[p1,np.ones(n)]@Ki.T)[:,:2].astype(np.float32)
    thr=1.5/600; t2=thr*thr
    R=T[:3,:3]; t=T[:3,3]; tx=np.array([[0,-t[2],t[1]],[t[2],0,-t[0]],[-t[1],t[0],0]]); Etrue=proj_E(tx@R)
    d,_=samp(Etrue,x.astype(np.float64),u.astype(np.float64)); print(n,frac,'true-model cost',np.minimum(d,t2).sum(),'recall',(d<t2)[inl].mean())
    for seed in (5,6):
        E,mask,ninl=find_essential_batch(torch.from_numpy(x)[None].cuda(),torch.from_numpy(u)[None].cuda(),None,thr=thr,iters=16384,seed=seed)
        Eg=E[0].double().cpu().numpy(); m=mask[0].cpu().numpy().astype(bool)
        d,_=samp(Eg,x.astype(np.float64),u.astype(np.float64))
        print('  seed',seed,'gpu cost',np.minimum(d,t2).sum(),'recall',m[inl].mean().round(3),'numpy-recall',(d<t2)[inl].mean().round(3),'svals',np.linalg.svd(Eg)[1].round(4))
        Ec=Eg
        for rd,mult in enumerate([3,3,2,2,1.5,1,1,1]):
            d,den=samp(Ec,x.astype(np.float64),u.astype(np.float64)); mm=d<t2*mult*mult
            Ec=fit(x[mm].astype(np.float64),u[mm].astype(np.float64),1/np.sqrt(den[mm]))
            d2,_=samp(Ec,x.astype(np.float64),u.astype(np.float64)); print('     numpy LO round',rd,'n',mm.sum(),'cost',np.minimum(d2,t2).sum().round(6),'recall',(d2<t2)[inl].mean().round(3))
