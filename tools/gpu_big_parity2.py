import sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner
n = int(sys.argv[1])
pos, faces = synth.grid_mesh(n, n, terrain=True)
mm = MeshMap(pos, faces); ed = mm.edgeDistances(); vc = np.zeros(mm.V, np.float32); mm.setCosts(vc, ed)
c = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
sf = int(2 * ((c // n) * (n - 1) + (c % n))); sp = pos[faces[sf]].mean(0).astype(np.float32)
om = O.OracleMesh(pos, faces); r = om.cvp(ed, vc, sf, sp)
mm.L.mnb_debug_set_sweeps.argtypes = [C.c_void_p, C.c_int32]
for (k, dl) in ((-1, 1.8), (0, 0.3)):
    mm.L.mnb_debug_set_sweeps(mm._ctx, k); mm.set_tuning(dl, 0, 0)
    g = CVPMeshPlanner(mm).waveFrontPropagation(sf, sp)
    bad = np.where(g['dist'].view(np.uint32) != r['dist'].view(np.uint32))[0]
    rel = np.abs(g['dist'][bad].astype(np.float64) - r['dist'][bad]) / r['dist'][bad]
    print(f"sweeps={k} delta={dl}: dist!= {bad.size} idx {bad[:8].tolist()} gpu {g['dist'][bad][:8].tolist()} ref {r['dist'][bad][:8].tolist()} maxrel {rel.max() if bad.size else 0:.3e} backsteps {r['backsteps']}", flush=True)
