import sys
sys.path.insert(0,'/root/repo/tools/exp')
import numpy as np, scipy.linalg as sla
from block_jacobi_sweeps import theta_like, block_jacobi
chi=int(sys.argv[1]); rng=np.random.default_rng(1)
th=theta_like(chi,rng)
sref=np.linalg.svd(th,compute_uv=False)
def rep(name,X):
    sw,s=block_jacobi(X,16,1)
    print("%-28s outer sweeps %2d  sigma err %.1e"%(name,sw,np.abs(s-sref).max()/sref[0]),flush=True)
rep("plain",th)
Q,R,P=sla.qr(th,pivoting=True,mode='economic')
rep("QRCP: jacobi on R^H",R.conj().T)
Q1,R1=np.linalg.qr(th)
rep("QR: jacobi on R^H",R1.conj().T)
Q2,R2=np.linalg.qr(R.conj().T)
rep("QRCP+QR: jacobi on R2^H",R2.conj().T)
rep("QRCP+QR: jacobi on R2",R2)
# sorted columns by norm
idx=np.argsort(-np.linalg.norm(th,axis=0)); rep("plain, columns sorted",th[:,idx])
