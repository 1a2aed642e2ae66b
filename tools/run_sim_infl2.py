import numpy as np, ctypes as C, sys
sys.path.insert(0,'.')
from oracle import oracle as O
from tests.util import *
L=C.CDLL('tools/libsimband.so'); vp=C.c_void_p
L.sim_inflation.argtypes=[C.c_uint32,C.c_uint32,vp,vp,C.c_uint32,vp,vp,vp,C.c_uint32,C.c_float,vp,vp,C.c_int]
def p(a): return None if a is None else a.ctypes.data_as(vp)
rng = np.random.default_rng(11)
pos, faces = mesh_case(150, True)
m = O.OracleMesh(pos, faces); ed = m.edge_distances()
le = np.unique(np.concatenate([disc_lethals(pos, 12, 0.35, seed=3), np.array([5, 5, 777, 12000], np.uint32)]))
invalid = (rng.random(m.V) < 0.01).astype(np.uint8)
kw = dict(inscribed_radius=0.5, inflation_radius=1.5, lethal_value=1.0, inscribed_value=0.9, cost_scaling_factor=1.0)
for inv in (None, invalid):
    ref = m.inflation(ed, le, invalid=inv, **kw)
    out=np.empty(m.V,np.float32); st=np.zeros(4)
    L.sim_inflation(m.V,m.F,p(m.faces),p(m.edges),m.E,p(ed),p(inv),p(le),le.size,1.5,p(out),p(st),-1)
    fr=np.isfinite(ref['dist']); fs=np.isfinite(out)
    both=fr&fs
    rel=np.abs(out[both]-ref['dist'][both])/np.maximum(ref['dist'][both],1e-30)
    print("inv",inv is not None,"rounds",st[0],"watchdog",st[2],"finite ref/sim",fr.sum(),fs.sum(),"set diff",(fr!=fs).sum(),"neq",(out[both]!=ref['dist'][both]).sum(),"maxrel",rel.max() if rel.size else 0, "pops", ref['pops'])
    bad=np.where(both&(out!=ref['dist']))[0]
    bad=bad[np.argsort(ref['dist'][bad])][:5]
    print("  first bad", [(int(b),float(ref['dist'][b]),float(out[b])) for b in bad])
