"""Dev script (GPU box): bit-parity of the whole-grid kernels vs the oracle on a mesh large enough for the adaptive band."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner, DijkstraMeshPlanner
n = int(sys.argv[1])
pos, faces = synth.grid_mesh(n, n, terrain=True)
mm = MeshMap(pos, faces); ed = mm.edgeDistances(); vc = np.zeros(mm.V, np.float32); mm.setCosts(vc, ed)
c = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
sf = int(2 * ((c // n) * (n - 1) + (c % n))); sp = pos[faces[sf]].mean(0).astype(np.float32)
g = CVPMeshPlanner(mm).waveFrontPropagation(sf, sp)
d = DijkstraMeshPlanner(mm).dijkstra(int(c))
om = O.OracleMesh(pos, faces)
t = time.time(); r = om.cvp(ed, vc, sf, sp); t1 = time.time() - t
print(f"n={n} cvp: gpu {g['kernel_ms']:.1f} ms rounds {g['rounds']} | oracle {r['seconds']:.2f} s | dist!= {(g['dist'].view(np.uint32) != r['dist'].view(np.uint32)).sum()} pred!= {(g['pred'] != r['pred']).sum()}", flush=True)
r = om.dijkstra(ed, vc, int(c))
print(f"n={n} dijkstra: gpu {d['kernel_ms']:.1f} ms rounds {d['rounds']} | oracle {r['seconds']:.2f} s | dist!= {(d['dist'].view(np.uint32) != r['dist'].view(np.uint32)).sum()} pred!= {(d['pred'] != r['pred']).sum()}", flush=True)
mm.close()
