"""Index-level emulation of invert_tile_mma (csrc/mas_assemble.cu, MAS_OPT_INVERT_VARIANT 4) in numpy: 8 warps x 32 lanes, the
m16n8k8 fragment layouts (A row-major: a0 (g,t) a1 (g+8,t) a2 (g,t+4) a3 (g+8,t+4); B "col": b0 (k=t,n=g) b1 (k=t+4,n=g);
C: c0 (g,2t) c1 (g,2t+1) c2 (g+8,2t) c3 (g+8,2t+1)), the half-tile ownership table and the shared-memory panels X / Y / S / W
exactly as the kernel addresses them, with exact arithmetic in place of 3xTF32.  Written while no GPU was available: it
checks that the staging, panel products, trailing updates and the reconstruction of E and D index the right elements
(the result must be the inverse to rounding).      python tools/emulate_mma_inversion.py
"""
import numpy as np
kPs=20; n=96
items=[[0x00,0x32,0x40,0x44,0x54,0xff],[0x01,0x33,0x41,0x45,0x55,0xff],[0x10,0x24,0x42,0x46,0x56,0xff],[0x11,0x25,0x43,0x47,0x57,0xff],[0x12,0x30,0x34,0x48,0x58,0xff],[0x13,0x31,0x35,0x49,0x59,0xff],[0x20,0x22,0x36,0x50,0x52,0x5a],[0x21,0x23,0x37,0x51,0x53,0x5b]]
rng=np.random.RandomState(0)
B=rng.randn(n,n); A=B@B.T+n*np.eye(n)
def mma(acc,a,b,lanes=range(32)):
    Am=np.zeros((16,8)); Bm=np.zeros((8,8))
    for lane in range(32):
        g,q=lane>>2,lane&3
        Am[g,q]=a[lane][0]; Am[g+8,q]=a[lane][1]; Am[g,q+4]=a[lane][2]; Am[g+8,q+4]=a[lane][3]
        Bm[q,g]=b[lane][0]; Bm[q+4,g]=b[lane][1]
    D=Am@Bm
    for lane in range(32):
        g,q=lane>>2,lane&3
        acc[lane][0]+=D[g,2*q]; acc[lane][1]+=D[g,2*q+1]; acc[lane][2]+=D[g+8,2*q]; acc[lane][3]+=D[g+8,2*q+1]
RL=lambda g,u: g+8*(u>>1)
CL=lambda th,q,u: 8*th+2*q+(u&1)
# acc[warp][e][lane][u]
acc=np.zeros((8,6,32,4))
def dec(w,e):
    it=items[w][e]; return it!=0xff, it>>4,(it>>1)&7,it&1
for w in range(8):
    for e in range(6):
        has,ti,tj,th=dec(w,e)
        if not has: continue
        for lane in range(32):
            g,q=lane>>2,lane&3
            for u in range(4): acc[w,e,lane,u]=A[16*ti+RL(g,u),16*tj+CL(th,q,u)]
X=np.full(96*kPs,np.nan); Y=np.full(96*kPs,np.nan); S=np.full(5*16*kPs,np.nan); W=np.full(16*kPs,np.nan)
def factor(Wsm):
    T=np.array([[Wsm[r*kPs+c] for c in range(16)] for r in range(16)])
    for x in range(15):
        for y in range(x+1,16):
            r=-T[y,x]/T[x,x]
            T[y,:]+=r*T[x,:]; T[y,x]=r   # (T[y,x] += r*T[x,x] then overwritten)
    d=np.diag(T).copy()
    Wq=np.zeros(16*kPs)
    for r in range(16):
        for c in range(16): Wq[r*kPs+c]= T[r,c] if c<r else (1.0 if c==r else 0.0)
    return Wq,d
for K in range(6):
    X[:]=X; 
    for w in range(8):
        for e in range(6):
            has,ti,tj,th=dec(w,e)
            if not has: continue
            for lane in range(32):
                g,q=lane>>2,lane&3
                for u in range(4):
                    rl,cl=RL(g,u),CL(th,q,u)
                    if ti==K and tj==K: W[rl*kPs+cl]=acc[w,e,lane,u]
                    elif ti==K: S[(tj*16+cl)*kPs+rl]=acc[w,e,lane,u]
                    elif tj==K: S[((ti-1)*16+rl)*kPs+cl]=acc[w,e,lane,u]
    Wq,dq=factor(W)
    for w in range(8):
        for e in range(6):
            has,ti,tj,th=dec(w,e)
            if not has: continue
            if ti==K and tj<K:
                o=np.zeros((32,4))
                for kk in range(2):
                    a=np.zeros((32,4)); b=np.zeros((32,2))
                    for lane in range(32):
                        g,q=lane>>2,lane&3; k0=8*kk+q
                        a[lane]=[Wq[g*kPs+k0],Wq[(g+8)*kPs+k0],Wq[g*kPs+k0+4],Wq[(g+8)*kPs+k0+4]]
                        Sj=(tj*16+8*th+g)*kPs
                        b[lane]=[S[Sj+k0],S[Sj+k0+4]]
                    mma(o,a,b)
                for lane in range(32):
                    g,q=lane>>2,lane&3
                    for u in range(4):
                        acc[w,e,lane,u]=o[lane][u]; Y[(tj*16+CL(th,q,u))*kPs+RL(g,u)]=o[lane][u]
            elif ti==K and tj==K:
                for lane in range(32):
                    g,q=lane>>2,lane&3
                    for u in range(4):
                        rl,cl=RL(g,u),CL(th,q,u); wv=Wq[rl*kPs+cl]
                        Y[(K*16+cl)*kPs+rl]=wv
                        acc[w,e,lane,u]= wv if rl>cl else (dq[rl] if rl==cl else 0.0)
            elif tj==K and ti>K:
                o=np.zeros((32,4)); Si=(ti-1)*16*kPs
                for kk in range(2):
                    a=np.zeros((32,4)); b=np.zeros((32,2))
                    for lane in range(32):
                        g,q=lane>>2,lane&3; k0=8*kk+q
                        a[lane]=[S[Si+g*kPs+k0],S[Si+(g+8)*kPs+k0],S[Si+g*kPs+k0+4],S[Si+(g+8)*kPs+k0+4]]
                        Wn=(8*th+g)*kPs
                        b[lane]=[Wq[Wn+k0],Wq[Wn+k0+4]]
                    mma(o,a,b)
                for lane in range(32):
                    g,q=lane>>2,lane&3
                    for u in range(4):
                        rl,cl=RL(g,u),CL(th,q,u)
                        Y[(ti*16+rl)*kPs+cl]=o[lane][u]; X[(ti*16+rl)*kPs+cl]=o[lane][u]/dq[cl]; acc[w,e,lane,u]=0.0
    if K==5: break
    for w in range(8):
        for e in range(6):
            has,ti,tj,th=dec(w,e)
            if not has or ti<=K: continue
            Xi=ti*16*kPs
            for kk in range(2):
                a=np.zeros((32,4)); b=np.zeros((32,2))
                for lane in range(32):
                    g,q=lane>>2,lane&3; k0=8*kk+q
                    Yj=(tj*16+8*th+g)*kPs
                    a[lane]=[-X[Xi+g*kPs+k0],-X[Xi+(g+8)*kPs+k0],-X[Xi+g*kPs+k0+4],-X[Xi+(g+8)*kPs+k0+4]]
                    b[lane]=[Y[Yj+k0],Y[Yj+k0+4]]
                o=[list(acc[w,e,l]) for l in range(32)]
                mma(o,a,b)
                acc[w,e]=np.array(o)
# reconstruct E (unit lower) and D
E=np.zeros((n,n)); D=np.zeros(n)
for w in range(8):
    for e in range(6):
        has,ti,tj,th=dec(w,e)
        if not has: continue
        for lane in range(32):
            g,q=lane>>2,lane&3
            for u in range(4):
                rl,cl=RL(g,u),CL(th,q,u); v=acc[w,e,lane,u]
                if ti==tj:
                    if rl==cl: D[16*ti+rl]=v
                    v= v if rl>cl else (1.0 if rl==cl else 0.0)
                E[16*ti+rl,16*tj+cl]=v
print("NaNs:",np.isnan(E).sum(),np.isnan(D).sum())
inv=E.T@np.diag(1/D)@E
print("||inv - A^-1|| rel:",np.abs(inv-np.linalg.inv(A)).max()/np.abs(np.linalg.inv(A)).max())
print("E A E^T - D:",np.abs(E@A@E.T-np.diag(D)).max())
