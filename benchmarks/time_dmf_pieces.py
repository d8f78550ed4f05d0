import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from adsurf_b200 import _capi
from adsurf_b200.workload import config5_ensemble
s = np.load('tests/golden/shipped_fixtures.npz'); obs = s['disper_02']
S=1024
d,a,b,rho = (torch.tensor(x, device='cuda') for x in config5_ensemble(S,20,0))
t = torch.tensor(np.tile(obs[:,0],(S,1)), device='cuda'); c = torch.tensor(np.tile(obs[:,1],(S,1)), device='cuda')
C = torch.tensor(np.tile(np.arange(0.5,5,0.01),(S,1)), device='cuda')
def tm(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("dmf_step none (points only) ms", tm(lambda: _capi.dmf_step(d,a,b,rho,t,c,None,_capi.COMPRESS_NONE)))
print("dmf_step exp (grid 450x300 + points) ms", tm(lambda: _capi.dmf_step(d,a,b,rho,t,c,C,_capi.COMPRESS_EXP)))
print("points_fwd ms", tm(lambda: _capi.points_fwd(d,a,b,rho,t,c)))
gF=torch.ones_like(c)
print("points_bwd ms", tm(lambda: _capi.points_bwd(d,a,b,rho,t,c,gF)))
print("grid_fwd minmax ms", tm(lambda: _capi.grid_fwd(d,a,b,rho,t,C,want_det=False,want_minmax=True)))
L=20
state={"b":b.clone(),"d":d.clone(),"alpha":a.clone(),"rho":rho.clone(),"mb":torch.zeros_like(b),"vb":torch.zeros_like(b),"md":torch.zeros_like(b),"vd":torch.zeros_like(b)}
consts={"b0":b.clone(),"alpha0":a.clone(),"rho0":rho.clone(),"lower_b":0.1*b,"upper_b":2.5*b,"mindepth":torch.full_like(b,0.01),"maxdepth":torch.full_like(b,50.0)}
out=_capi.dmf_step(d,a,b,rho,t,c,None,_capi.COMPRESS_NONE)
loss=out[0]; grads=out[3:]
it=torch.zeros((S,L),dtype=torch.float64,device='cuda')
print("ensemble_adam_step ms", tm(lambda: _capi.ensemble_adam_step(state,grads,loss,consts,vp_mode=1,vp_vs_ratio=2.45,invert_thick=False,layering=0,damp_vertical=0.0,damp_horizontal=0.0,lr=0.03,beta1=0.9,beta2=0.999,eps=1e-8,step=3,iter_b=it,iter_d=None)))
print("ensemble_adam_step thick+damp ms", tm(lambda: _capi.ensemble_adam_step(state,grads,loss,consts,vp_mode=1,vp_vs_ratio=2.45,invert_thick=True,layering=2,damp_vertical=0.01,damp_horizontal=0.01,lr=0.03,beta1=0.9,beta2=0.999,eps=1e-8,step=3,iter_b=it,iter_d=it)))
