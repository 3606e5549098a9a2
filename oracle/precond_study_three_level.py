#!/usr/bin/env python
"""Design study, full three-level emulation (TEST INFRASTRUCTURE / evidence, not product
code): outer PCG on the condensed system with M^-1 = D^-1 + P Ac^-1 P^T, where the inner
coarse solve (relative residual 1e-2) is itself preconditioned by Jacobi + a piecewise-
constant aggregation level over k x k ELEMENT TILES (a vertex joins the tile of the first
element containing it; dense inverse at the top).  Uses the product's host tables.

    python oracle/precond_study_three_level.py

Output on the development container:
    n=32  p=4 tiles 4x4   (64 dofs):  two-level 23 outer / 1026 inner; three-level 23 outer / 250 inner
    n=64  p=4 tiles 8x8   (64 dofs):  two-level 23 outer / 2012 inner; three-level 23 outer / 403 inner
    n=128 p=2 tiles 8x8   (256 dofs): two-level 18 outer / 2769 inner; three-level 18 outer / 284 inner
    n=128 p=2 tiles 16x16 (64 dofs):  two-level 18 outer / 2769 inner; three-level 18 outer / 470 inner
The outer count is unchanged and the inner count per outer iteration (16-26) no longer grows
with the mesh.
"""
import sys, numpy as np, os
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import sem_oracle as so
from scipy import sparse
from scipy.sparse.linalg import splu
from spectralelementmethod_b200.condensed import condensed_tables, coarse_tables

def pcg(apply, b, M, rtol, x0=None, maxiter=100000):
    x = np.zeros_like(b) if x0 is None else x0.copy(); r = b-apply(x); z=M(r); p=z.copy(); rz=r@z; bb=b@b; it=0
    while it<maxiter and r@r > rtol*rtol*bb:
        Ap=apply(p); a=rz/(p@Ap); x+=a*p; r-=a*Ap; z=M(r); rzn=r@z; p=z+(rzn/rz)*p; rz=rzn; it+=1
    return x,it

def run(n,p,k):
    basis=so.Basis(p); N=p+1; NE=4*p
    nodes,l2g=so.build_case("C",n,n,p,True,False)
    geo=so.geometry(basis,nodes,l2g)
    c=so.condensed_system(p,geo["invJ"],geo["JxW"],l2g)
    on,vals=so.dirichlet_data(nodes,l2g,geo["x_phys"],so.mesh_boundary_faces(n,n))
    n_ext=c["n_ext"]; ids=c["ids"]; S_e=c["S"]; D=on[:n_ext]
    l2g_ext,nptr,npos=condensed_tables(np.pad(ids,((0,0),(0,N*N-NE))).astype(np.uint32),np.arange(NE),n_ext)
    nptr=nptr.astype(np.int64)
    ct=coarse_tables(l2g_ext,nptr,npos,D,basis.nodes)
    vc=ct["vert_c"].astype(np.int64); Dc=ct["dirichlet_c"]; nv=ct["n_v"]
    Phi=ct["phi"][None]*(~D)[ids][:,:,None]*(~Dc)[vc][:,None,:]
    Ace=np.einsum("eka,ekj,ejc->eac",Phi,S_e,Phi)
    Ac=sparse.coo_matrix((Ace.reshape(-1),(np.repeat(vc,4,axis=1).ravel(),np.tile(vc,(1,4)).ravel())),shape=(nv,nv)).tocsr()+sparse.diags(Dc.astype(float))
    dc=Ac.diagonal()
    # aggregates from ELEMENT tiles (k x k elements): a vertex joins the tile of the first element containing it
    ex,ey=np.divmod(np.arange(n*n),n); tile=(ex//k)*((n+k-1)//k)+(ey//k)
    first=ct["vpos"].astype(np.int64)[ct["vptr"].astype(np.int64)[:-1]]//4
    agg=tile[first]; uniq,agg=np.unique(agg,return_inverse=True)
    P2=sparse.coo_matrix((np.where(Dc,0.0,1.0),(np.arange(nv),agg)),shape=(nv,uniq.size)).tocsr()
    P2=P2[:,np.asarray(P2.sum(axis=0)).ravel()>0]
    A3=(P2.T@Ac@P2).toarray(); A3inv=np.linalg.inv(A3)
    S=c["Sg"]; Mf=(~D).astype(float)
    fine=lambda u: Mf*(S@(Mf*u))+(1-Mf)*u
    sd=np.where(D,1.0,S.diagonal()); 
    pv,pw=ct["pv"].astype(np.int64),ct["pw"]; rptr,ridx,rw=ct["rptr"].astype(np.int64),ct["ridx"].astype(np.int64),ct["rw"]
    def restrict(r):
        out=np.zeros(nv); nz=rptr[1:]>rptr[:-1]; out[nz]=np.add.reduceat(rw*r[ridx],rptr[:-1][nz]); return out
    prolong=lambda xc: pw[:,0]*xc[pv[:,0]]+pw[:,1]*xc[pv[:,1]]
    inner=[]
    def M3(q): return q/dc + P2@(A3inv@(P2.T@q))
    def M(r, three):
        rc=restrict(r)
        xc,itc=pcg(lambda v:Ac@v, rc, (M3 if three else (lambda q:q/dc)), 1e-2); inner.append(itc)
        return r/sd+prolong(xc)
    gv=np.where(D,vals[:n_ext],0.0); b=c["grhs"]-Mf*(S@gv); b[D]=gv[D]; x0=np.where(D,b,0.0)
    inner.clear(); x2,it2=pcg(fine,b,lambda r:M(r,False),1e-12,x0); in2=sum(inner)
    inner.clear(); x3,it3=pcg(fine,b,lambda r:M(r,True),1e-12,x0); in3=sum(inner)
    print("n=%d p=%d tiles %dx%d (third level %d dofs): two-level %d outer / %d inner;  three-level %d outer / %d inner;  diff %.1e"%(n,p,k,k,A3.shape[0],it2,in2,it3,in3,np.linalg.norm(x3-x2)/np.linalg.norm(x2)),flush=True)
run(32,4,4); run(64,4,8); run(128,2,8); run(128,2,16)
