import numpy as np, ctypes as C, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
from run_sim2 import *
n=140; rng = np.random.default_rng(5)
pos, faces = mesh_case(n, True)
m = O.OracleMesh(pos, faces); ed = m.edge_distances()
vc = np.where(rng.random(m.V) < 0.04, 1.2, rng.random(m.V) * 0.7).astype(np.float32)
v, f, sp = centre_seed(pos, faces, (0.3,0.35))
for x in faces[f]: vc[x] = 0.1
w = m.edge_weights(vc, ed, 1.0)
pop = np.full(m.V, 0xffffffff, np.uint32)
O.lib().orc_debug_set_pop_buffer(pop.ctypes.data_as(C.c_void_p))
ref = m.cvp(w, vc, f, sp)
O.lib().orc_debug_set_pop_buffer(None)
for c in (7876, 7735, 7736, 7877, 7734, 7594,7593):
    print(c, "d", ref['dist'][c], "pop#", pop[c], "pred", ref['pred'][c], "cut", ref['cutting_face'][c], faces[ref['cutting_face'][c]])
order = np.argsort(pop)
i = int(pop[7735])
print("pops around 7735:", [(int(x), float(ref['dist'][x])) for x in order[i-6:i+4]])
for fc in (15360, 15363):
    vs=list(faces[fc]); c=7876; k=vs.index(c); v1=vs[(k+1)%3]; v2=vs[(k+2)%3]
    d=ref['dist'].copy(); d[c]=np.inf
    pred=np.arange(m.V,dtype=np.uint32); dr=np.zeros(m.V,np.float32); cut=-np.ones(m.V,np.int32)
    ok=m.cvp_wavefront_update(w,int(fc),int(v1),int(v2),c,d,pred,dr,cut)
    e=lambda a,b: w[[i for i,e in enumerate(m.edges.tolist()) if e==[min(a,b),max(a,b)]][0]]
    print("face",fc,"v1",v1,"v2",v2,"->",ok,d[c],"weights c,b,a:",e(v1,v2),e(v1,c),e(v2,c), "euclid", np.linalg.norm(pos[v1]-pos[v2]),np.linalg.norm(pos[v1]-pos[c]),np.linalg.norm(pos[v2]-pos[c]), "costs", vc[v1],vc[v2],vc[c])
