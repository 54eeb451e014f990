"""Why the opt-in tensor-core sweep of the 96 / 128-variable shapes (option sweep_dmma) sits 1.7e-7 N from qpOASES where the DFMA\nsweep sits 4e-9 N: scalar Gauss-Jordan against BLOCK Gauss-Jordan (explicitly inverted 8 x 8 pivot blocks, with and without the\nkernels' D - I trick) in numpy float64 on the mixed-gait h = 16 Hessians, errors against an extended-precision inverse.\nTest infrastructure (uses oracle/cmpc_numpy.py)."""
import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'quad-periodic-mpc_b200'))
from oracle import cmpc_numpy as N
from cmpc_b200 import synth
h=16
inst = synth.make_batch(64, horizon=h, seed=77, gaits=("trot","bound","pace","gallop"), n_segment=10, spread=1.5)
def true_inv(H):
    Hl=H.astype(np.longdouble); K=np.linalg.inv(H).astype(np.longdouble)
    for _ in range(3): K = K + K@(np.eye(len(H),dtype=np.longdouble) - Hl@K)
    return K
def scalar_gj(A):
    A=A.copy(); n=len(A)
    for p in range(n):
        d=1.0/A[p,p]; col=A[:,p].copy(); row=A[p,:].copy()
        A=A-np.outer(col*d,row); A[:,p]=-col*d; A[p,:]=row*d; A[p,p]=d   # in-place GJ -> inverse (sign conventions: standard)
    return A
def inv8(D):
    return scalar_gj(D)
def block_gj_trick(A, bs=8):
    # the kernel's scheme: panel C~ = rows of block s with (D - I) in the diagonal block; M = -D^-1 C~; A += C~' M; final K = -(A - 2 I)
    A=A.copy(); n=len(A); nb=(n+bs-1)//bs
    for s in range(nb):
        lo,hi=s*bs,min(n,(s+1)*bs)
        D=A[lo:hi,lo:hi].copy(); C=A[lo:hi,:].copy(); C[:,lo:hi]=D-np.eye(hi-lo)
        M=-(inv8(D)@C)
        A=A+C.T@M
    return -(A-2*np.eye(n))
def block_gj_plain(A, bs=8):
    # textbook block Gauss-Jordan (explicit pivot-block inverse, no D - I trick)
    A=A.copy(); n=len(A); nb=(n+bs-1)//bs
    for s in range(nb):
        lo,hi=s*bs,min(n,(s+1)*bs); P=slice(lo,hi)
        Di=inv8(A[P,P]); rows=A[P,:].copy(); cols=A[:,P].copy()
        A=A-cols@(Di@rows)
        A[P,:]=Di@rows; A[:,P]=-cols@Di; A[P,P]=Di
    return A
res=[]
for idx in range(24):
    H,g=N.condense_closed(inst, idx); keep=N.contact_vars(inst["gait"][idx],h)
    H=H[np.ix_(keep,keep)]; g=g[keep]; n=len(g)
    if n<70: continue
    e=int(np.floor(np.log2(np.diag(H).max())))+1; sc=2.0**(-e); Hs=H*sc
    Kt=true_inv(Hs)
    x_t=-(Kt@g.astype(np.longdouble))
    out=[n, float(np.linalg.cond(Hs))]
    for f in (scalar_gj, block_gj_plain, block_gj_trick):
        K=f(Hs)
        errK=float(np.abs(K-Kt).max()/np.abs(Kt).max())
        x=-(K@g)*1.0
        errx=float(np.abs(x*sc - (x_t*sc)).max())
        out+= [errK, errx]
    res.append(out)
    print("n=%3d cond %.1e | scalar GJ: relK %.1e dx0 %.1e | block GJ: relK %.1e dx0 %.1e | block GJ with D-I trick: relK %.1e dx0 %.1e"%tuple(out))
r=np.array(res)
print("median dx0: scalar %.1e  block %.1e  block+trick %.1e"%(np.median(r[:,3]),np.median(r[:,5]),np.median(r[:,7])))
