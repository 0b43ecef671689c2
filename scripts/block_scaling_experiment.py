#!/usr/bin/env python
"""Host experiment (numpy / scipy, oracle): 9 x 9 block scaling S~ = B^-1 S B^-T with B_c = chol(U_c + lambda I) against the
point-Jacobi scaling the device uses, as the matrix the FP32 factor of the mixed solver is computed from: condition
number of the scaled reduced camera system and FP64 CG iterations with an FP32 Cholesky factor (exact spotrf, and with a
relative perturbation of 3.5e-7 per entry, the backward error measured for the tensor-core factorisation).
profiles/r02_block_scaling_experiment.md.    python scripts/block_scaling_experiment.py trafalgar-257"""
import sys, os, numpy as np, scipy.linalg as sla
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
import bundleadjustment.jl_b200.synth as synth
from oracle import oracle as O
O.build()
from test_mixed_prototype import _schur_system
def cg(S,b,M,tol=1e-13,maxit=60):
    x=np.zeros_like(b); r=b.copy(); z=M(r); p=z.copy(); rz=r@z; bn=np.linalg.norm(b)
    for it in range(1,maxit+1):
        q=S@p; a=rz/(p@q); x+=a*p; r-=a*q
        if np.linalg.norm(r)/bn<=tol: return it
        z=M(r); rzn=r@z; p=z+(rzn/rz)*p; rz=rzn
    return maxit
def chol32(St, eps, rng):
    E = rng.normal(size=St.shape)*eps*np.abs(St); E=(E+E.T)/2
    try: return sla.cholesky((St+E).astype(np.float32), lower=True).astype(np.float64)
    except Exception: return None
shape=sys.argv[1] if len(sys.argv)>1 else "trafalgar-257"
p=synth.make_problem(shape)
vals=O.jac_coord(p.cam_idx,p.pnt_idx,p.x0,p.npnts,8).reshape(-1,2,12)
Bk=vals[:,:,3:]
U=np.zeros((p.ncams,9,9)); np.add.at(U,p.cam_idx-1,np.einsum('kia,kib->kab',Bk,Bk))
rng=np.random.default_rng(0)
for lam in (39.0,4.0,0.5,0.05,0.006,6e-4):
    S,b=_schur_system(O,p,lam,nt=8)
    n=S.shape[0]; nc=n//9
    d=np.sqrt(np.diag(U.reshape(nc,9,9))[...] if False else np.concatenate([np.diag(U[c])+lam for c in range(nc)]))
    St=S/d[:,None]/d[None,:]
    Binv=np.zeros((n,n))
    for c in range(nc):
        Lc=np.linalg.cholesky(U[c]+lam*np.eye(9)); Binv[9*c:9*c+9,9*c:9*c+9]=np.linalg.inv(Lc)
    St2=Binv@S@Binv.T
    out=[lam]
    for name,Sx in (("jacobi(U+lam)",St),("block chol(U+lam)",St2)):
        w=np.linalg.eigvalsh(Sx); cond=w[-1]/w[0]
        its=[]
        for eps in (0.0,3.5e-7):
            L=chol32(Sx,eps,rng)
            if L is None: its.append(-1); continue
            if name.startswith("jacobi"):
                M=lambda r,L=L: sla.solve_triangular(L.T, sla.solve_triangular(L, r/d, lower=True), lower=False)/d
            else:
                M=lambda r,L=L: Binv.T@sla.solve_triangular(L.T, sla.solve_triangular(L, Binv@r, lower=True), lower=False)
            its.append(cg(S,b,M))
        out.append((name,"cond %.2e"%cond,"max|S~| %.3f"%np.abs(Sx).max(),its))
    print(out,flush=True)
