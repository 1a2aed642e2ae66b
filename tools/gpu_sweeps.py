"""Dev script (GPU box): in-round sweeps of the whole-grid CVP kernel -- rounds / time / bit-equality vs sweeps = 0."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner

def run(n, combos, weighted=False):
    pos, faces = synth.grid_mesh(n, n, terrain=True)
    mm = MeshMap(pos, faces)
    ed = mm.edgeDistances()
    if weighted:
        vc = (0.45 + 0.45 * np.sin(3.0 * pos[:, 0]) * np.cos(2.0 * pos[:, 1])).astype(np.float32)
        w = mm.computeEdgeWeights(vc, 1.0)
    else:
        vc = np.zeros(mm.V, np.float32); w = ed
    mm.setCosts(vc, w)
    c = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
    sf = int(np.where((faces == c).any(1))[0][0]); sp = pos[faces[sf]].mean(0).astype(np.float32)
    pl = CVPMeshPlanner(mm)
    mm.L.mnb_debug_set_sweeps.argtypes = [C.c_void_p, C.c_int32]
    ref = None
    for (k, delta) in combos:
        mm.L.mnb_debug_set_sweeps(mm._ctx, k); mm.set_tuning(delta, 0, 0)
        best = 1e9
        for it in range(2):
            g = pl.waveFrontPropagation(sf, sp); best = min(best, g['kernel_ms'])
        if ref is None: ref = g['dist'].copy()
        ne = int((g['dist'].view(np.uint32) != ref.view(np.uint32)).sum())
        print(f"n={n} weighted={int(weighted)} sweeps={k} delta={delta}: kernel_ms={best:.2f} rounds={g['rounds']} recomp/V={g['recomputes']/mm.V:.2f} "
              f"dist!=ref {ne} settled={g['settled']}", flush=True)
    mm.close()

if __name__ == "__main__":
    n = int(sys.argv[1]); weighted = len(sys.argv) > 3 and sys.argv[3] == 'w'
    combos = [tuple(float(x) if i else int(x) for i, x in enumerate(c.split(':'))) for c in sys.argv[2].split(',')]
    run(n, combos, weighted)
