import numpy as np, ctypes as C, sys, time
sys.path.insert(0,'.')
from oracle import oracle as O
from tests.util import *
L=C.CDLL('tools/libsimband.so'); vp=C.c_void_p
L.sim_inflation.argtypes=[C.c_uint32,C.c_uint32,vp,vp,C.c_uint32,vp,vp,vp,C.c_uint32,C.c_float,vp,vp,C.c_int]
def p(a): return None if a is None else a.ctypes.data_as(vp)
n=int(sys.argv[1])
pos, faces = mesh_case(n, True)
m = O.OracleMesh(pos, faces); ed = m.edge_distances()
t=time.time(); lay=m.layers(); print("layers",time.time()-t)
le=np.union1d(np.where(lay['lethal_mask']!=0)[0], disc_lethals(pos, 200, 0.3)).astype(np.uint32)
t=time.time(); ref = m.inflation(ed, le); print("oracle inflation", time.time()-t, "lethals", le.size, "pops", ref['pops'])
out=np.empty(m.V,np.float32); st=np.zeros(4)
t=time.time(); L.sim_inflation(m.V,m.F,p(m.faces),p(m.edges),m.E,p(ed),None,p(le),le.size,0.4,p(out),p(st),-1); print("sim",time.time()-t)
fr=np.isfinite(ref['dist']); fs=np.isfinite(out); both=fr&fs
print("rounds",st[0],"watchdog",st[2],"finite ref/sim",fr.sum(),fs.sum(),"neq",(out[both]!=ref['dist'][both]).sum())
