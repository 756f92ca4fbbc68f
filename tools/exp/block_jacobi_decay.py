import sys
sys.path.insert(0,'/root/repo/tools/exp')
import numpy as np
from block_jacobi_sweeps import theta_like, inner_jacobi
def run(A, b, tol=1.6e-14, max_sweeps=40):
    A=A.copy(); n=A.shape[1]; nb=n//b; sref=np.linalg.svd(A,compute_uv=False)
    for sw in range(max_sweeps):
        ring=list(range(nb)); mx=0.0; act=0
        for _ in range(nb-1):
            for k in range(nb//2):
                i,j=ring[k],ring[nb-1-k]
                cols=np.r_[i*b:(i+1)*b, j*b:(j+1)*b]
                P=A[:,cols]; G=P.conj().T@P
                d=np.sqrt(np.abs(np.diag(G).real)); off=np.abs(G)-np.diag(np.abs(np.diag(G)))
                r=(off/np.outer(d,d)).max(); mx=max(mx,r)
                if r<=tol: continue
                act+=1
                W,_=inner_jacobi(G,1,tol); A[:,cols]=P@W
            ring=[ring[0]]+[ring[-1]]+ring[1:-1]
        s=np.sort(np.linalg.norm(A,axis=0))[::-1]
        print("sweep %2d: max scaled offdiag seen %.2e, active pairs %4d, sigma err after sweep %.1e"%(sw+1,mx,act,np.abs(s-sref).max()/sref[0]),flush=True)
        if act==0: break
rng=np.random.default_rng(1)
run(theta_like(int(sys.argv[1]),rng),16)
