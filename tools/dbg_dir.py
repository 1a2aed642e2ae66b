import sys, numpy as np
sys.path.insert(0,'.')
from oracle import oracle as O
from mesh_navigation_b200 import api as A
from tests.util import *
pos, faces = mesh_case(100, False)
om = O.OracleMesh(pos, faces); mm = A.MeshMap(pos, faces); mm.set_tuning(0.3,1,0)
ed = om.edge_distances(); vc = np.zeros(om.V, np.float32); mm.setCosts(vc, ed)
v, f, sp = centre_seed(pos, faces)
ref = om.cvp(ed, vc, f, sp); got = A.CVPMeshPlanner(mm).waveFrontPropagation(f, sp)
bad = np.where(np.abs(got['direction']-ref['direction'])>1e-6)[0]
print(len(bad))
for c in bad[:6]:
    fc = ref['cutting_face'][c]; vs=list(faces[fc]); k=vs.index(c); v1=vs[(k+1)%3]; v2=vs[(k+2)%3]
    d=ref['dist'].copy(); d[c]=np.inf; pred=np.arange(om.V,dtype=np.uint32); dr=np.zeros(om.V,np.float32); cut=-np.ones(om.V,np.int32)
    om.cvp_wavefront_update(ed,int(fc),int(v1),int(v2),int(c),d,pred,dr,cut)
    print(c,"got dir",got['direction'][c],"ref dir",ref['direction'][c],"re-eval dir",dr[c],"pred",got['pred'][c],ref['pred'][c],pred[c],"cut",got['cutting_face'][c],fc,"d",got['dist'][c],ref['dist'][c],d[c])
