import sys, numpy as np
sys.path.insert(0,'.')
from oracle import oracle as O
from mesh_navigation_b200 import api as A
from tests.util import *
pos, faces = mesh_case(100, False)
om = O.OracleMesh(pos, faces); mm = A.MeshMap(pos, faces)
ed = om.edge_distances()
le = disc_lethals(pos, 8, 0.3)
ref = om.inflation(ed, le); got = A.InflationLayer(mm).waveCostInflation(le)
fr=np.isfinite(ref['dist'])
bad=np.where(fr & (got['dist']!=ref['dist']))[0]
print("lethals",le.size,"labelled",fr.sum(),"mismatch",bad.size,"rounds",got['rounds'], "max ref dist", ref['dist'][fr].max())
order=bad[np.argsort(ref['dist'][bad])]
for c in order[:4]:
    print("v",c,"ref",ref['dist'][c],"got",got['dist'][c])
    for fc in np.where((faces==c).any(1))[0]:
        vs=list(faces[fc]); k=vs.index(c); v1=vs[(k+1)%3]; v2=vs[(k+2)%3]
        e=lambda a,b: ed[[i for i,e in enumerate(om.edges.tolist()) if e==[min(a,b),max(a,b)]][0]]
        a_,b_,c_=e(v2,c),e(v1,c),e(v1,v2)
        dot=np.float32((a_*a_+b_*b_-c_*c_)/(2*a_*b_))
        cand=O.lib().orc_sethian_update(float(ref['dist'][v1]),float(ref['dist'][v2]),float(a_),float(b_),float(dot),1.0)
        print("   face",fc,"v1",v1,ref['dist'][v1],got['dist'][v1],"v2",v2,ref['dist'][v2],got['dist'][v2],"cand(ref)",cand)
import ctypes as C
lab=np.zeros((om.V,4),np.uint32)
mm.L.mnb_debug_get_labels.argtypes=[C.c_void_p,C.c_void_p]
mm.L.mnb_debug_get_labels(mm._ctx, lab.ctypes.data_as(C.c_void_p))
f=lab.view(np.float32)
for v in (3731,3732,3733,3831,3832,3833,3834,3932,3933):
    print(v,"d",f[v,0],"a1",f[v,1],"a2",f[v,2],"a3w",hex(lab[v,3]),"ref",ref['dist'][v])
k=lab[3733,1]
same=np.where(lab[:,1]==k)[0]
print("vertices with a1==",f[3733,1],":",same)
for v in same:
    print(" ",v,"d",f[v,0],"a1",f[v,1],"a2",f[v,2],"a3w",hex(lab[v,3]),"ref",ref['dist'][v], "lethal", v in set(le.tolist()))
    for fc in np.where((faces==v).any(1))[0]:
        vs=list(faces[fc]); kk=vs.index(v); v1=vs[(kk+1)%3]; v2=vs[(kk+2)%3]
        print("      face",fc,"v1",v1,f[v1,0],f[v1,1],"v2",v2,f[v2,0],f[v2,1])
