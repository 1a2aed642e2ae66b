import numpy as np, ctypes as C, sys
sys.path.insert(0,'.')
from oracle import oracle as O
from tests.util import *
L=C.CDLL('tools/libsimband.so'); vp=C.c_void_p
L.sim_inflation.argtypes=[C.c_uint32,C.c_uint32,vp,vp,C.c_uint32,vp,vp,vp,C.c_uint32,C.c_float,vp,vp,C.c_int]
def p(a): return None if a is None else a.ctypes.data_as(vp)
pos, faces = mesh_case(100, False)
m = O.OracleMesh(pos, faces); ed = m.edge_distances()
le = np.unique(disc_lethals(pos, 8, 0.3))
ref = m.inflation(ed, le)
out=np.empty(m.V,np.float32); st=np.zeros(4)
L.sim_inflation(m.V,m.F,p(m.faces),p(m.edges),m.E,p(ed),None,p(le),le.size,0.4,p(out),p(st),int(sys.argv[1]) if len(sys.argv)>1 else -1)
fr=np.isfinite(ref['dist'])
print("rounds",st[0],"watchdog",st[2],"finite ref/sim",fr.sum(),np.isfinite(out).sum(),"neq",(out[fr]!=ref['dist'][fr]).sum())
bad=np.where(fr&(out!=ref['dist']))[0]
print(bad[:10], [(float(ref['dist'][b]),float(out[b])) for b in bad[:6]])
