import numpy as np, ctypes as C, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
from run_sim import *
n=int(sys.argv[1]); terrain=int(sys.argv[2]); fl=int(sys.argv[3]); dl=float(sys.argv[4])
pos,faces=synth.grid_mesh(n,n,terrain=bool(terrain))
m=O.OracleMesh(pos,faces); ed=m.edge_distances(); vc=np.zeros(m.V,np.float32)
seed=synth.nearest_vertex(pos,[n*0.05,n*0.05,pos[:,2].mean()])
sf=int(np.where((faces==seed).any(1))[0][0]); sp=pos[faces[sf]].mean(0).astype(np.float32)
ref=m.cvp(ed,vc,sf,sp)
out=np.empty(m.V,np.float32); st=np.zeros(4)
L.sim_cvp_band(m.V,m.F,p(m.faces),p(m.edges),m.E,p(m.pos),p(ed),p(vc),None,sf,p(sp),1.0,dl,fl,p(out),p(st))
diff=np.where(out!=ref['dist'])[0]
print("ndiff",diff.size)
order=diff[np.argsort(ref['dist'][diff])]
for v in order[:3]:
    print("vertex",v,"ref",ref['dist'][v],"sim",out[v],"pred",ref['pred'][v],"cut",ref['cutting_face'][v])
    fs=np.where((faces==v).any(1))[0]
    for f in fs:
        vs=faces[f]; print("   face",f,vs,[ (float(ref['dist'][x]),float(out[x])) for x in vs])
