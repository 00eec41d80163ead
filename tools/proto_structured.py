"""Numpy prototype of the structured condensation used by the CUDA kernel (dev aid; validated against the oracle).

Blocks (nDOF = 1 shown; layout [u | q=(i,delta) | l=(f,a)]):
  W = M^-1,  Squ_d[i][j] := S_qu[(i,d), j],  Suq_d[i][j] := S_uq[i, (j,d)]
  per face f (t x t): LL_f, LU_f [(f,a), fn(b)], UL_f [fn(a),(f,b)], LQ_fd [(f,a),(fn(b),d)], QL_fd [(fn(a),d),(f,b)]
"""
import numpy as np, sys
sys.path.insert(0, '.')
from oracle.refel import ReferenceElement
from oracle import lib as O

def structured(re, A, F):
    d, nN, nNf, nFc = re.dim, re.nNodes, re.faceElement.nNodes, re.nFaces
    u, q, l = nN, nN*d, nFc*nNf
    fn = np.array(re.faceNodes)
    sQ, sL = u, u+q
    Suu = A[:u,:u]; 
    Squ = [A[sQ+np.arange(nN)*d+k, :u] for k in range(d)]          # [d] (nN x nN)
    Suq = [A[:u, sQ+np.arange(nN)*d+k] for k in range(d)]
    M = A[sQ+np.arange(nN)*d][:, sQ+np.arange(nN)*d]
    W = np.linalg.inv(M)
    LL = [A[sL+f*nNf:sL+(f+1)*nNf, sL+f*nNf:sL+(f+1)*nNf] for f in range(nFc)]
    LU = [A[sL+f*nNf:sL+(f+1)*nNf][:, fn[f]] for f in range(nFc)]
    UL = [A[fn[f]][:, sL+f*nNf:sL+(f+1)*nNf] for f in range(nFc)]
    LQ = [[A[sL+f*nNf:sL+(f+1)*nNf][:, sQ+fn[f]*d+k] for k in range(d)] for f in range(nFc)]
    QL = [[A[sQ+fn[f]*d+k][:, sL+f*nNf:sL+(f+1)*nNf] for k in range(d)] for f in range(nFc)]
    # check the sparsity assumptions
    chk = A.copy()
    chk[:u,:u]=0; chk[sQ:sL,:u]=0; chk[:u,sQ:sL]=0
    for k in range(d):
        idx = sQ+np.arange(nN)*d+k; chk[np.ix_(idx,idx)]=0
    for f in range(nFc):
        sl = slice(sL+f*nNf, sL+(f+1)*nNf)
        chk[sl,sl]=0; chk[sl, fn[f]]=0; chk[fn[f], sl]=0
        for k in range(d):
            chk[sl, sQ+fn[f]*d+k]=0; chk[sQ+fn[f]*d+k, sl]=0
    assert np.abs(chk).max()==0, np.abs(chk).max()
    Ad = [W@Squ[k] for k in range(d)]
    B = [np.zeros((nN,l)) for k in range(d)]
    for k in range(d):
        for f in range(nFc):
            B[k][:, f*nNf:(f+1)*nNf] = W[:, fn[f]] @ QL[f][k]
    K = Suu - sum(Suq[k]@Ad[k] for k in range(d))
    R = np.zeros((u,l))
    for f in range(nFc):
        R[fn[f], f*nNf:(f+1)*nNf] += UL[f]
    R -= sum(Suq[k]@B[k] for k in range(d))
    Ki = np.linalg.inv(K)
    U = -Ki@R; U0 = Ki@F[:u]
    Qd = [-(Ad[k]@U) - B[k] for k in range(d)]
    Q0d = [-(Ad[k]@U0) for k in range(d)]
    S = np.zeros((l,l)); S0 = F[sL:].copy()
    for f in range(nFc):
        sl = slice(f*nNf,(f+1)*nNf)
        S[sl] = LU[f]@U[fn[f]] + sum(LQ[f][k]@Qd[k][fn[f]] for k in range(d))
        S[sl, sl] += LL[f]
        S0[sl] -= LU[f]@U0[fn[f]] + sum(LQ[f][k]@Q0d[k][fn[f]] for k in range(d))
    Q = np.zeros((q,l)); Q0=np.zeros(q)
    for k in range(d):
        Q[np.arange(nN)*d+k] = Qd[k]; Q0[np.arange(nN)*d+k]=Q0d[k]
    return U,Q,S,U0,Q0,S0

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for dim in (2,3):
        for p in (1,2,3):
            re = ReferenceElement(dim,p); rc = O.RefElC(re)
            nodes = re.nodes*0.3 + 0.02*rng.standard_normal(re.nodes.shape)   # curved, small
            nFc,nNf,nN = rc.nFc, rc.nNf, rc.nN
            tau = 1+rng.random(nFc*nNf); vel = rng.standard_normal((nN,dim)); diff = 0.5+rng.random(nN)
            md = O.make_model(1, O.OP_DIFFUSION|O.OP_CONVECTION|O.OP_REACTION|O.OP_SOURCE, diffComps=1, timeScheme=O.TS_EULER_IMPLICIT, dt=0.1)
            A,F = O.local_system(rc, md, nodes=nodes, tau=tau, diff=diff, vel=vel, srcIP=rng.random(rc.nIP), reacIP=rng.random(rc.nIP), solOld=rng.random(nN))
            u,q,l,n = O.sizes(rc,1)
            ref = O.condense(u,q,l,A,F,0); lu = O.condense(u,q,l,A,F,1)
            st = structured(re,A,F)
            for name,a,b,c in zip("U Q S U0 Q0 S0".split(), ref, lu, st):
                sc = np.abs(a).max()
                print(dim,p,name, "QRvsLU %.2e"%(np.abs(a-b).max()/sc), "QRvsStruct %.2e"%(np.abs(a-c).max()/sc))
