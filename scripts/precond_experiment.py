import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0,'/root/repo')
from oracle import oracle as O

def morton(ix,iy,r):
    m=0
    for b in range(r): m |= ((ix>>b)&1)<<(2*b) | ((iy>>b)&1)<<(2*b+1)
    return m

def interior_system(l, cor, c):
    rowptr,col,val,F = O.assemble(l, cor, c)
    N=O.n_dofs(l); n=1<<l
    K=sp.csr_matrix((val,col.astype(np.int64),rowptr.astype(np.int64)),shape=(N,N))
    d=O.dof_map(l)               # [jy,jx] -> dof
    perm=d.ravel()               # lex -> dof
    Kl=K[perm][:,perm]           # lexicographic
    jy,jx=np.meshgrid(np.arange(n+1),np.arange(n+1),indexing='ij')
    inter=((jx>0)&(jx<n)&(jy>0)&(jy<n)).ravel()
    A=Kl[inter][:,inter].tocsr()
    # rhs for basis 0: g = bilinear
    s=jx/n; t=jy/n
    g=((1-s)*(1-t)).ravel()
    b=-(Kl[inter][:,~inter] @ g[~inter])
    return A,b,n

def prolong(nc):  # interior coarse (nc-1)^2 -> interior fine (2nc-1)^2, bilinear
    nf=2*nc
    def P1(nc):
        rows=[];cols=[];vals=[]
        for I in range(1,nc):
            i=2*I
            for di,w in ((-1,0.5),(0,1.0),(1,0.5)):
                rows.append(i+di-1); cols.append(I-1); vals.append(w)
        return sp.csr_matrix((vals,(rows,cols)),shape=(nf-1,nc-1))
    p=P1(nc)
    return sp.kron(p,p).tocsr()

def pcg(A,b,M,tol=1e-12,maxit=5000):
    x=np.zeros_like(b); r=b.copy(); z=M(r); p=z.copy(); rz=r@z; it=0
    if np.linalg.norm(r)<=tol: return x,0
    while it<maxit:
        it+=1
        q=A@p; al=rz/(p@q); x+=al*p; r-=al*q
        if np.linalg.norm(r)<=tol: break
        z=M(r); rz2=r@z; p=z+(rz2/rz)*p; rz=rz2
    return x,it

def build_levels(A,n):
    Ls=[A]; Ps=[]
    nc=n
    while nc>2:
        nc//=2
        P=prolong(nc); Ps.append(P); Ls.append((P.T@Ls[-1]@P).tocsr())
    return Ls,Ps

def make_bpx(Ls,Ps,exact_coarsest=True):
    Ds=[1.0/L.diagonal() for L in Ls]
    Ac=Ls[-1].toarray(); Aci=np.linalg.inv(Ac)
    def M(r):
        rs=[r]
        for P in Ps: rs.append(P.T@rs[-1])
        zs=[Ds[i]*rs[i] for i in range(len(rs))]
        if exact_coarsest: zs[-1]=Aci@rs[-1]
        z=zs[-1]
        for i in range(len(Ps)-1,-1,-1):
            z=zs[i]+Ps[i]@z
        return z
    return M

def make_vcycle(Ls,Ps,om=0.8,nu=1):
    Ds=[1.0/L.diagonal() for L in Ls]
    Aci=np.linalg.inv(Ls[-1].toarray())
    def V(lv,r):
        if lv==len(Ls)-1: return Aci@r
        A=Ls[lv]; e=np.zeros_like(r)
        for _ in range(nu): e=e+om*Ds[lv]*(r-A@e)
        d=r-A@e
        e=e+Ps[lv]@V(lv+1,Ps[lv].T@d)
        for _ in range(nu): e=e+om*Ds[lv]*(r-A@e)
        return e
    return lambda r: V(0,r)

def make_twolevel(Ls,Ps,lvl,om=1.0):
    # additive: D^-1 + P_lvl Ac^-1 P_lvl^T
    D0=1.0/Ls[0].diagonal()
    Aci=np.linalg.inv(Ls[lvl].toarray())
    def M(r):
        rc=r
        for P in Ps[:lvl]: rc=P.T@rc
        zc=Aci@rc
        for P in Ps[:lvl][::-1]: zc=P@zc
        return om*D0*r+zc
    return M

def make_ssor(A,om=1.6, colour=False, n=None):
    if colour:
        m=n-1
        jy,jx=np.meshgrid(np.arange(m),np.arange(m),indexing='ij')
        colr=((jx%2)+2*(jy%2)).ravel()
        perm=np.argsort(colr,kind='stable')
        Ap=A[perm][:,perm].tocsr()
    else:
        perm=np.arange(A.shape[0]); Ap=A
    D=Ap.diagonal(); L=sp.tril(Ap,-1).tocsr(); U=sp.triu(Ap,1).tocsr()
    DL=(sp.diags(D)+om*L).tocsr(); DU=(sp.diags(D)+om*U).tocsr()
    inv=np.empty_like(perm); inv[perm]=np.arange(len(perm))
    def M(r):
        rp=r[perm]
        y=spla.spsolve_triangular(DL,rp,lower=True)
        y=D*y
        z=spla.spsolve_triangular(DU,y,lower=False)
        return (om*(2-om)*z)[inv]
    return M

cases=[("target per", 8,6,(77,200),O.COEFF_PERIODIC,(1/64,0.9999),0),
       ("target ref", 8,6,(9,3),O.COEFF_REFERENCE,(),0),
       ("cfg3",7,6,(100,17),O.COEFF_PERIODIC,(1/64,0.9999),0),
       ("cfg4 incl",8,5,(200,31),O.COEFF_INCLUSIONS,(2.0**-11,0.2,1e4,1.0),1234),
       ("cfg2",5,5,(5,7),O.COEFF_PERIODIC,(1/64,0.9999),0),
       ("default128",3,7,(3,5),O.COEFF_REFERENCE,(),0)]
for name,r,l,(ix,iy),kind,par,seed in cases:
    cor=O.coarse_corners(r,[morton(ix,iy,r)])[0]
    A,b,n=interior_system(l,cor,O.coeff(kind,par,seed))
    Ls,Ps=build_levels(A,n)
    D0=1.0/A.diagonal()
    res={}
    res['jacobi']=pcg(A,b,lambda r:D0*r)[1]
    res['ssor']=pcg(A,b,make_ssor(A))[1]
    res['ssor4c']=pcg(A,b,make_ssor(A,1.6,True,n))[1]
    res['ssor4c_w1.2']=pcg(A,b,make_ssor(A,1.2,True,n))[1]
    res['bpx']=pcg(A,b,make_bpx(Ls,Ps))[1]
    res['V11_j.8']=pcg(A,b,make_vcycle(Ls,Ps,0.8,1))[1]
    res['V22_j.8']=pcg(A,b,make_vcycle(Ls,Ps,0.8,2))[1]
    res['2lvl_H8']=pcg(A,b,make_twolevel(Ls,Ps,3))[1]
    res['2lvl_H4']=pcg(A,b,make_twolevel(Ls,Ps,2))[1]
    print(name, res, flush=True)

print("---- BPX variants ----")
def bpx_generic(Ls,Ps,nlev=None,exact=False,wts=None):
    nl=len(Ls) if nlev is None else nlev
    Ds=[1.0/L.diagonal() for L in Ls]
    Aci=np.linalg.inv(Ls[nl-1].toarray()) if exact else None
    def M(r):
        rs=[r]
        for P in Ps[:nl-1]: rs.append(P.T@rs[-1])
        zs=[Ds[i]*rs[i]*(1.0 if wts is None else wts[i]) for i in range(nl)]
        if exact: zs[-1]=Aci@rs[-1]
        z=zs[-1]
        for i in range(nl-2,-1,-1):
            z=zs[i]+Ps[i]@z
        return z
    return M
for name,r,l,(ix,iy),kind,par,seed in cases:
    cor=O.coarse_corners(r,[morton(ix,iy,r)])[0]
    A,b,n=interior_system(l,cor,O.coeff(kind,par,seed))
    d=A.diagonal(); S=sp.diags(1/np.sqrt(d))
    Ah=(S@A@S).tocsr(); bh=S@b
    Ls,Ps=build_levels(A,n); Lh,Ph=build_levels(Ah,n)
    res={}
    res['bpx']=pcg(A,b,bpx_generic(Ls,Ps))[1]
    res['bpx_scaled']=pcg(Ah,bh,bpx_generic(Lh,Ph),tol=1e-12/np.sqrt(d.max()))[1]
    for nl in (3,4,5):
        if nl<len(Ls):
            res['bpx_%dlev'%nl]=pcg(A,b,bpx_generic(Ls,Ps,nl))[1]
            res['bpx_%dlev_ex'%nl]=pcg(A,b,bpx_generic(Ls,Ps,nl,True))[1]
    print(name,res,flush=True)

print("---- hybrid variants ----")
def make_hybrid(Ls,Ps,om=0.8,coarse_scale=1.0,post=True):
    A=Ls[0]; Ds=[1.0/L.diagonal() for L in Ls]
    def C(d):
        rs=[d]
        for P in Ps: rs.append(P.T@rs[-1])
        z=Ds[-1]*rs[-1]
        for i in range(len(Ps)-1,0,-1):
            z=Ds[i]*rs[i]+Ps[i]@z
        return Ps[0]@z
    def M(r):
        e=om*Ds[0]*r
        d=r-A@e
        e=e+coarse_scale*C(d)
        if post: e=e+om*Ds[0]*(r-A@e)
        return e
    return M
for name,r,l,(ix,iy),kind,par,seed in cases:
    cor=O.coarse_corners(r,[morton(ix,iy,r)])[0]
    A,b,n=interior_system(l,cor,O.coeff(kind,par,seed))
    Ls,Ps=build_levels(A,n)
    res={}
    res['bpx']=pcg(A,b,bpx_generic(Ls,Ps))[1]
    for om in (0.7,0.9):
        for cs in (0.5,1.0):
            res['hyb_om%.1f_cs%.1f'%(om,cs)]=pcg(A,b,make_hybrid(Ls,Ps,om,cs))[1]
    res['V11']=pcg(A,b,make_vcycle(Ls,Ps,0.8,1))[1]
    print(name,res,flush=True)
