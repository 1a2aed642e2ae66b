import numpy as np, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
from run_sim import *
n=1000
pos,faces=synth.grid_mesh(n,n,terrain=False)
m=O.OracleMesh(pos,faces); ed=m.edge_distances(); vc=np.zeros(m.V,np.float32)
seed=synth.nearest_vertex(pos,[n*0.05,n*0.05,pos[:,2].mean()])
sf=int(np.where((faces==seed).any(1))[0][0]); sp=pos[faces[sf]].mean(0).astype(np.float32)
ref=m.cvp(ed,vc,sf,sp)
v=int(sys.argv[1])
fs=np.where((faces==v).any(1))[0]
for f in fs:
    vs=list(faces[f]); k=vs.index(v); v1=vs[(k+1)%3]; v2=vs[(k+2)%3]
    d=ref['dist'].copy(); d[v]=np.inf
    pred=np.arange(m.V,dtype=np.uint32); dr=np.zeros(m.V,np.float32); cut=-np.ones(m.V,np.int32)
    ok=m.cvp_wavefront_update(ed,int(f),int(v1),int(v2),v,d,pred,dr,cut)
    print(f, "v1",v1,ref['dist'][v1],"v2",v2,ref['dist'][v2],"T",max(ref['dist'][v1],ref['dist'][v2]),"->",ok,d[v],pred[v],dr[v])
