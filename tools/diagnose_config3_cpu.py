#!/usr/bin/env python
"""Config 3 diagnosis on the CPU (oracle matrices + scipy): the unsteady aSIMPLE composition of the reference
(lab_new/src/NSSolver.hpp:294-350) inside FGMRES(30), once with EXACT solves for F and S = B diag(F)^-1 Bt (sparse LU) and once with
the reference's single ILU(0) applications, on the three kinds of systems a time step meets: the first_iter branch (Stokes operator
without the 1/dt mass, nu = 1), the Newton branch at the first Reynolds stage (nu = 1) and at the last one (nu = 1/91).
usage: diagnose_config3_cpu.py gmsh|NX,NY MAX_IT      (results: profiles/r02_config3_diagnosis.md)"""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import nsxlib as N

def fgmres(A, b, M, tol, maxit, restart=30):
    x = np.zeros_like(b); it = 0; hist=[]
    while True:
        r = b - A(x); beta = np.linalg.norm(r); hist.append(beta)
        if beta <= tol or it >= maxit: return x, it, hist
        V=[r/beta]; Z=[]; H=np.zeros((restart+1,restart)); 
        for j in range(restart):
            z = M(V[j]); Z.append(z); w = A(z)
            for i in range(j+1):
                H[i,j]=w@V[i]; w-=H[i,j]*V[i]
            H[j+1,j]=np.linalg.norm(w); V.append(w/H[j+1,j]); it+=1
            e1=np.zeros(j+2); e1[0]=beta
            y,res,_,_=np.linalg.lstsq(H[:j+2,:j+1],e1,rcond=None)
            rn=np.linalg.norm(H[:j+2,:j+1]@y-e1); hist.append(rn)
            if rn<=tol or it>=maxit: break
        x = x + sum(y[i]*Z[i] for i in range(len(y)))
        if rn<=tol or it>=maxit:
            return x, it, hist

which = sys.argv[1]
d = N.Disc.from_gmsh(N.golden_mesh_path()) if which=='gmsh' else N.Disc.generate(*[int(v) for v in which.split(',')], triangles=True)
o = N.Oracle(d, inlet_amplitude=0.3)
print("dofs", d.n, "n_u", d.n_u, "n_p", d.n_p)
for mode, nu, label in ((N.MODE_UNSTEADY_FIRST, 1.0, "first_iter nu=1"), (N.MODE_UNSTEADY_NEWTON, 1.0, "newton nu=1"), (N.MODE_UNSTEADY_NEWTON, 1/91., "newton nu=1/91")):
    o.vec(0)[:]=0; o.vec(1)[:]=0
    if mode==N.MODE_UNSTEADY_NEWTON:
        s=N.synthetic_state(d,5,noise=1e-3)*3; o.vec(0)[:]=s; o.vec(1)[:]=s
    r0=o.assemble(mode, mode==N.MODE_UNSTEADY_FIRST, nu, 0.01)
    F,Bt,B = o.csr(N.BLOCK_F).tocsc(), o.csr(N.BLOCK_BT).tocsr(), o.csr(N.BLOCK_B).tocsr()
    J = o.jacobian().tocsr(); b=o.vec(3).copy()
    D = F.diagonal(); S = (B@sp.diags(1/D)@Bt).tocsc()
    nu_, np_ = d.n_u, d.n_p
    alpha=0.5
    for kind in ("exact","ilu0"):
        if kind=="exact":
            Fs=spl.splu(F); Ss=spl.splu(S); fF=Fs.solve; fS=Ss.solve
        else:
            # ILU(0) via oracle inner_apply on F; S needs the schur in the oracle
            o.schur()
            fF=lambda x:o.inner_apply(N.BLOCK_F,1,x); fS=lambda x:o.inner_apply(N.BLOCK_S,1,x)
        def M(src):
            su,sp_=src[:nu_],src[nu_:]
            ut=fF(su); t=sp_+B@ut; p=fS(t)/alpha
            return np.concatenate([ut-(Bt@p)/D, p])
        t=time.time()
        x,it,hist=fgmres(lambda v:J@v, b, M, 1e-6, 60 if kind=="exact" else int(sys.argv[2]))
        print(f"{label:18s} {kind:6s}: its {it:6d} res0 {hist[0]:.3e} final {hist[-1]:.3e}  ({time.time()-t:.1f}s) hist every 300: {[f'{h:.1e}' for h in hist[::300]][:12]}")
