"""Dev script (GPU box): Dijkstra whole-grid kernel timing / parity vs the oracle."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, DijkstraMeshPlanner

def run(n, combos, check=True):
    pos, faces = synth.grid_mesh(n, n, terrain=True)
    mm = MeshMap(pos, faces)
    ed = mm.edgeDistances(); vc = np.zeros(mm.V, np.float32); mm.setCosts(vc, ed)
    seed = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
    mm.L.mnb_debug_set_sweeps.argtypes = [C.c_void_p, C.c_int32]
    ref = None
    if check:
        om = O.OracleMesh(pos, faces); ref = om.dijkstra(ed, vc, seed)
    pl = DijkstraMeshPlanner(mm)
    for (k, delta, cl) in combos:
        mm.L.mnb_debug_set_sweeps(mm._ctx, int(k)); mm.set_tuning(delta, int(cl), 0)
        best = 1e9
        for it in range(2):
            g = pl.dijkstra(seed); best = min(best, g['kernel_ms'])
        msg = ""
        if ref is not None:
            msg = f" dist!= {(g['dist'].view(np.uint32) != ref['dist'].view(np.uint32)).sum()} pred!= {(g['pred'] != ref['pred']).sum()} cpu={ref['seconds']*1e3:.0f}ms"
        print(f"n={n} sweeps={k} delta={delta} cluster={cl}: kernel_ms={best:.2f} rounds={g['rounds']} recomp/V={g['recomputes']/mm.V:.2f}{msg}", flush=True)
    mm.close()

if __name__ == "__main__":
    n = int(sys.argv[1])
    combos = [tuple(float(x) for x in c.split(':')) for c in sys.argv[2].split(',')]
    run(n, combos, check=len(sys.argv) < 4)
