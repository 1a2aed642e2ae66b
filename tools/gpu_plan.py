"""Dev script (GPU box): time the pieces of a makePlan (locate / wavefront with robot face / back-tracking)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner
n = int(sys.argv[1])
pos, faces = synth.grid_mesh(n, n, terrain=True)
mm = MeshMap(pos, faces); ed = mm.edgeDistances(); mm.setCosts(np.zeros(mm.V, np.float32), ed)
corner = lambda u, v: pos[int(v * (n - 1)) * n + int(u * (n - 1))]
pts = np.stack([corner(0.1, 0.1), corner(0.9, 0.9)]).astype(np.float32)
pl = CVPMeshPlanner(mm)
for rep in range(3):
    t0 = time.perf_counter(); nv, fc, ba = mm.locate(pts); t1 = time.perf_counter(); ls = mm.stats()
    rc = mm.L.mnb_cvp(mm._ctx, int(fc[1]), pts[1].ctypes.data, int(fc[0]), 1.0, 0.3, None, None, None, None); t2 = time.perf_counter(); ws = mm.stats()
    bt = pl.backtrack(pts[0], int(fc[0])); t3 = time.perf_counter()
    print(f"locate {1e3*(t1-t0):.2f} ms (kernel {ls['kernel_ms']:.3f}) | wavefront {1e3*(t2-t1):.2f} ms (kernel {ws['kernel_ms']:.2f}, rounds {ws['rounds']}) rc={rc} | "
          f"backtrack {1e3*(t3-t2):.2f} ms (kernel {bt['kernel_ms']:.2f}) points {len(bt['positions'])} outcome {bt['outcome']}", flush=True)
mm.close()
