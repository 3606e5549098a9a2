#!/usr/bin/env python
"""Design study for a THIRD level under the vertex coarse space (TEST INFRASTRUCTURE /
evidence, not product code): iterations of the inner coarse solve (relative residual 1e-2)
with Jacobi alone and with Jacobi + a piecewise-constant aggregation level over k x k
vertex blocks (exact solve there).

    python oracle/precond_study_levels.py

Output on the development container (order 2 elements, curved cells; the coarse operator
barely depends on the order):
    n=64  n_v=4225 : Jacobi-PCG 57 its;  + aggregation (k, its, dofs): (4, 12, 289) (8, 19, 81) (16, 30, 25)
    n=128 n_v=16641: Jacobi-PCG 171 its; + aggregation: (4, 12, 1089) (8, 20, 289) (16, 34, 81)
    n=256 n_v=66049: Jacobi-PCG 241 its; + aggregation: (4, 12, 4225) (8, 20, 1089) (16, 35, 289)
i.e. the inner iteration count stops growing with the mesh (20 at k = 8) where Jacobi-PCG
needs ~n; at config 2 (n = 1024) the measured 25 119 inner iterations per solve would
drop to ~32 x 20.
"""
import sys, numpy as np
import os; HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import sem_oracle as so
from scipy import sparse
from scipy.sparse.linalg import splu
from spectralelementmethod_b200.condensed import condensed_tables, coarse_tables

def pcg(A, b, M, rtol, maxiter=100000):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r@z; bb=b@b; it=0
    while it<maxiter and r@r > rtol*rtol*bb:
        Ap=A@p; a=rz/(p@Ap); x+=a*p; r-=a*Ap; z=M(r); rzn=r@z; p=z+(rzn/rz)*p; rz=rzn; it+=1
    return x,it

p=2
for n in (64,128,256):
    basis=so.Basis(p); N=p+1; NE=4*p
    nodes,l2g=so.build_case("C",n,n,p,True,False)
    geo=so.geometry(basis,nodes,l2g)
    c=so.condensed_system(p,geo["invJ"],geo["JxW"],l2g)
    on,vals=so.dirichlet_data(nodes,l2g,geo["x_phys"],so.mesh_boundary_faces(n,n))
    n_ext=c["n_ext"]; ids=c["ids"]; S_e=c["S"]; D=on[:n_ext]
    l2g_ext,nptr,npos=condensed_tables(np.pad(ids,((0,0),(0,N*N-NE))).astype(np.uint32),np.arange(NE),n_ext)
    ct=coarse_tables(l2g_ext,nptr,npos,D,basis.nodes)
    vc=ct["vert_c"].astype(np.int64); Dc=ct["dirichlet_c"]; nv=ct["n_v"]
    Phi=ct["phi"][None]*(~D)[ids][:,:,None]*(~Dc)[vc][:,None,:]
    Ace=np.einsum("eka,ekj,ejc->eac",Phi,S_e,Phi)
    rows=np.repeat(vc,4,axis=1).ravel(); cols=np.tile(vc,(1,4)).ravel()
    Ac=sparse.coo_matrix((Ace.reshape(-1),(rows,cols)),shape=(nv,nv)).tocsr()
    Ac=Ac+sparse.diags(Dc.astype(float))
    dc=Ac.diagonal()
    rng=np.random.default_rng(0); b=rng.standard_normal(nv); b[Dc]=0
    _,itj=pcg(Ac,b,lambda r:r/dc,1e-2)
    # vertex grid coords: vertices are (n+1)x(n+1) in compact order sorted by exterior id -> recover (i,j) from coordinates
    vids=np.unique(ids[:,:4])
    out=[]
    for k in (4,8,16):
        # aggregate by element blocks of k x k: vertex (i,j) -> aggregate (i//k, j//k) using lattice index from coordinates rank
        # use lattice coordinates: exterior ids are sorted by old lexicographic id => vertices in lexicographic (i,j) order
        ii,jj=np.divmod(np.arange(nv),n+1)
        agg=(ii//k)*((n+k)//k+1)+(jj//k)
        uniq,agg=np.unique(agg,return_inverse=True)
        P2=sparse.coo_matrix((np.where(Dc,0.0,1.0),(np.arange(nv),agg)),shape=(nv,uniq.size)).tocsr()
        keep=np.asarray(P2.sum(axis=0)).ravel()>0
        P2=P2[:,keep]
        A2=(P2.T@Ac@P2).tocsc(); lu=splu(A2)
        _,it3=pcg(Ac,b,lambda r:r/dc+P2@lu.solve(P2.T@r),1e-2)
        out.append((k,it3,A2.shape[0]))
    print("n=%d n_v=%d: inner Jacobi-PCG to 1e-2: %d its; + aggregation level (k, its, dofs): %s"%(n,nv,itj,out),flush=True)
