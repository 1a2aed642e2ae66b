"""Dev script (GPU box): full-field single CVP plan on the 5M terrain -- the 8-lane whole-grid kernel (k_cvp_grid) against the
lean batch round loop run by the whole grid (k_cvp_batch<0>, MNB_GRID_ENGINE=1) at several band widths."""
import sys, zlib, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2240
pos, faces = synth.grid_mesh(n, n, terrain=True)
mm = MeshMap(pos, faces)
ed = mm.edgeDistances(); mm.setCosts(np.zeros(mm.V, np.float32), ed)
c = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
sf = int(2 * ((c // n) * (n - 1) + (c % n))); sp = pos[faces[sf]].mean(0).astype(np.float32)
mm.L.mnb_debug_set_grid_engine.argtypes = [C.c_void_p, C.c_int32, C.c_float]
pl = CVPMeshPlanner(mm)
def run(tag):
    best = None
    for rep in range(3):
        g = pl.waveFrontPropagation(sf, sp)
        best = g["kernel_ms"] if best is None else min(best, g["kernel_ms"])
    print(f"{tag}: kernel_ms={best:.2f} rounds={g['rounds']} evals/V={g['recomputes']/mm.V:.2f} skipped/V={g.get('skipped',0)/mm.V:.2f} crc={zlib.crc32(g['dist'].tobytes()):08x} pred_crc={zlib.crc32(g['pred'].tobytes()):08x}", flush=True)
run("k_cvp_grid (8 lanes)")
for dw in [float(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else "1.5,2.5,4,8,16".split(','))]:
    mm.L.mnb_debug_set_grid_engine(mm._ctx, 1, dw)
    run(f"k_cvp_batch<0> delta={dw}w")
mm.close()
