"""Lane-level NumPy emulation of `potrf_diag_mma_kernel` (csrc/condvar.cu): the 16-column panel factorisation
(warp-0 register Cholesky + inverse with shuffles, DMMA panel and trailing update on 8 x 8 fragments) and the blocked
forward-substitution inverse, with the SAME index expressions as the CUDA code and `mma.sync.m8n8k4.f64` modelled
fragment by fragment (a: row = lane / 4, k = lane % 4; b: k = lane % 4, n = lane / 4; c: row = lane / 4,
cols 2 (lane % 4) + {0, 1}).  Used on the CPU of the build container (no GPU there) to validate the index arithmetic
before the first GPU run, and to reproduce the RBF blocks of BASELINE configs[0] on which a block-doubling inverse
lost the positive pivots of the NEXT block.

    python tools/emulate_potrf_diag.py          # one ill-conditioned 1,024 x 1,024 RBF matrix, 8 diagonal blocks (~6 min)
"""
import numpy as np

NB = 128; PA = 132; PT = 68; PD = 20
def diag_emul(S):
    global A,T,invd,Dinv
    A=np.zeros(NB*PA); T=np.zeros(64*PT); invd=np.zeros(NB); Dinv=np.zeros(16*PD)
    for r in range(NB):
        for c in range(NB):
            A[r*PA+c]=S[r,c] if c<=r else 0.0
    def inv_at(r,c):
        return A[c*PA+r] if r>c else (invd[r] if r==c else 0.0)
    def dmma(acc, a_l, b_l):
        # acc: dict lane->[c0,c1]; a_l, b_l: per-lane scalars (arrays of 32)
        Am=np.zeros((8,4)); Bm=np.zeros((4,8))
        for lane in range(32):
            g,t=lane>>2,lane&3
            Am[g,t]=a_l[lane]; Bm[t,g]=b_l[lane]
        C=Am@Bm
        for lane in range(32):
            g,t=lane>>2,lane&3
            acc[lane][0]+=C[g,2*t]; acc[lane][1]+=C[g,2*t+1]
    def factor16(c0):
        a=np.zeros((32,16)); myinv=np.zeros(32)
        for lane in range(32):
            r=lane&15
            for k in range(16): a[lane,k]=A[(c0+r)*PA+c0+k]
        for k in range(16):
            akk=a[k,k]  # shfl from lane k
            d=np.sqrt(akk); idv=1.0/d
            lrk=np.zeros(32)
            for lane in range(32):
                r=lane&15
                lrk[lane]=d if r==k else a[lane,k]*idv
                a[lane,k]=lrk[lane]
                if r==k: myinv[lane]=idv
            for c in range(k+1,16):
                lck=lrk[c]
                for lane in range(32):
                    a[lane,c]=a[lane,c]-lrk[lane]*lck
        x=np.zeros((32,16))
        for lane in range(32):
            c=lane&15
            x[lane,c]=myinv[lane]
        for rr in range(1,16):
            s=np.zeros(32)
            for k in range(rr):
                lrk=a[rr,k]
                for lane in range(32): s[lane]+=lrk*x[lane,k]
            irr=myinv[rr]
            for lane in range(32):
                c=lane&15
                if rr>c: x[lane,rr]=-s[lane]*irr
        for lane in range(16):
            r=lane; c=lane
            for k in range(16):
                if k<=r: A[(c0+r)*PA+c0+k]=a[lane,k]
                else: A[(c0+c)*PA+c0+k]=x[lane,k]
                Dinv[k*PD+c]=x[lane,k]
            invd[c0+r]=myinv[lane]
    for c0 in range(0,NB,16):
        factor16(c0)
        nfr=(NB-c0-16)//8
        for warp in range(8):
            for f in range(warp,nfr,8):
                rr=c0+16+8*f
                av=np.zeros((4,32))
                for ks in range(4):
                    for lane in range(32):
                        g,t=lane>>2,lane&3
                        av[ks,lane]=A[(rr+g)*PA+c0+ks*4+t]
                acc=[[{l:[0.0,0.0] for l in range(32)}][0] for nf in range(2)]
                acc=[{l:[0.0,0.0] for l in range(32)} for nf in range(2)]
                for ks in range(4):
                    for nf in range(2):
                        b=np.array([Dinv[(nf*8+(l>>2))*PD+ks*4+(l&3)] for l in range(32)])
                        dmma(acc[nf],av[ks],b)
                for nf in range(2):
                    for lane in range(32):
                        g,t=lane>>2,lane&3
                        A[(rr+g)*PA+c0+nf*8+t*2]=acc[nf][lane][0]
                        A[(rr+g)*PA+c0+nf*8+t*2+1]=acc[nf][lane][1]
        for warp in range(8):
            for fi in range(warp,nfr,8):
                ri=c0+16+8*fi
                av=np.zeros((4,32))
                for ks in range(4):
                    for lane in range(32):
                        g,t=lane>>2,lane&3
                        av[ks,lane]=A[(ri+g)*PA+c0+ks*4+t]
                for fk in range(fi+1):
                    rk=c0+16+8*fk
                    u={l:[0.0,0.0] for l in range(32)}
                    for ks in range(4):
                        b=np.array([A[(rk+(l>>2))*PA+c0+ks*4+(l&3)] for l in range(32)])
                        dmma(u,av[ks],b)
                    for lane in range(32):
                        g,t=lane>>2,lane&3
                        q=(ri+g)*PA+rk+t*2
                        A[q]-=u[lane][0]; A[q+1]-=u[lane][1]
    PR=116
    for bi in range(1,NB//16):
        r0=bi*16; nfr=r0//8
        for warp in range(8):
            for idx in range(warp,2*nfr,8):
                mf=idx&1; nf=idx>>1
                u={l:[0.0,0.0] for l in range(32)}
                for ks in range(2*nf,r0//4):
                    a=np.array([A[(r0+mf*8+(l>>2))*PA+ks*4+(l&3)] for l in range(32)])
                    bb=np.array([inv_at(ks*4+(l&3), nf*8+(l>>2)) for l in range(32)])
                    dmma(u,a,bb)
                for lane in range(32):
                    g,t=lane>>2,lane&3
                    T[(mf*8+g)*PR+nf*8+t*2]=u[lane][0]; T[(mf*8+g)*PR+nf*8+t*2+1]=u[lane][1]
        for n in range(r0):
            x=np.zeros(16)
            for r in range(16):
                sacc=-T[r*PR+n]
                for k in range(r):
                    sacc=sacc-A[(r0+r)*PA+r0+k]*x[k]
                x[r]=sacc*invd[r0+r]
            for r in range(16):
                A[n*PA+r0+r]=x[r]
    L=np.zeros((NB,NB)); Li=np.zeros((NB,NB))
    for r in range(NB):
        for c in range(NB):
            if c<=r: L[r,c]=A[r*PA+c]
            Li[r,c]=inv_at(r,c)
    return L,Li

def potrf(S,diag):
    S=S.copy(); m=S.shape[0]; nb=m//NB
    for j in range(nb):
        d=slice(j*NB,(j+1)*NB)
        L,X=diag(S[d,d]); S[d,d]=L
        if np.isnan(L).any(): print("NaN in block",j); return S
        print("block",j,"ok min diag",np.diag(L).min(),flush=True)
        if j+1<nb:
            S[(j+1)*NB:,d]=S[(j+1)*NB:,d]@X.T
            P=S[(j+1)*NB:,d]
            S[(j+1)*NB:,(j+1)*NB:]-=P@P.T
    return np.tril(S)
x=np.linspace(-2,8,2000)[:,None]
K=np.exp(-0.5*(x-x.T)**2)
if __name__ == "__main__":
  S=K[:1024,:1024]+1e-10*np.eye(1024)
  with np.errstate(all="ignore"):
      L=potrf(S,diag_emul)
  print("nan",np.isnan(L).any(), "residual", np.abs(L@L.T-S).max())
