import numpy as np, ctypes as C, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
from run_sim2 import *
n=300; fac=1.0; rng = np.random.default_rng(5)
pos, faces = mesh_case(n, True)
m = O.OracleMesh(pos, faces); ed = m.edge_distances()
vc = np.where(rng.random(m.V) < 0.04, 1.2, rng.random(m.V) * 0.7).astype(np.float32)
v, f, sp = centre_seed(pos, faces, (0.3,0.35))
for x in faces[f]: vc[x] = 0.1
w = m.edge_weights(vc, ed, fac)
pop = np.full(m.V, 0xffffffff, np.uint32)
O.lib().orc_debug_set_pop_buffer(pop.ctypes.data_as(C.c_void_p))
ref = m.cvp(w, vc, f, sp)
O.lib().orc_debug_set_pop_buffer(None)
out = np.empty(m.V, np.float32); st = np.zeros(4)
L.sim_cvp_band(m.V, m.F, p(m.faces), p(m.edges), m.E, p(m.pos), p(w), p(vc), None, f, p(sp), 1.0, 0.3, 0, p(out), p(st))
fin=np.isfinite(ref['dist'])
idx = np.where(fin)[0][np.argsort(pop[fin])]
bad = [c for c in idx if out[c]!=ref['dist'][c]][:2]
for c in bad:
    print("first diff (by pop order) v",c,"ref",ref['dist'][c],"sim",out[c],"pop#",pop[c],"pred",ref['pred'][c],"cut",ref['cutting_face'][c])
    for fc in np.where((faces==c).any(1))[0]:
        vs=list(faces[fc]); k=vs.index(c); v1=vs[(k+1)%3]; v2=vs[(k+2)%3]
        d=ref['dist'].copy(); d[c]=np.inf
        pred=np.arange(m.V,dtype=np.uint32); dr=np.zeros(m.V,np.float32); cut=-np.ones(m.V,np.int32)
        ok=m.cvp_wavefront_update(w,int(fc),int(v1),int(v2),int(c),d,pred,dr,cut)
        print("    face",fc,"v1",v1,(float(ref['dist'][v1]),float(out[v1]),int(pop[v1])),"v2",v2,(float(ref['dist'][v2]),float(out[v2]),int(pop[v2])),"cand(ref inputs)",ok,d[c])
L.sim_get_label.argtypes=[C.c_uint32, C.c_void_p]
for x in (22174, 22474, 22775, 22475, 22776, 22175):
    o=np.zeros(4); L.sim_get_label(x, o.ctypes.data_as(C.c_void_p))
    print(x, "sim label d,tau,sig,minor", o, "2x", 2*x, "ref d", ref['dist'][x], "pop", pop[x], "refpred", ref['pred'][x], "cut", ref['cutting_face'][x], faces[ref['cutting_face'][x]])
