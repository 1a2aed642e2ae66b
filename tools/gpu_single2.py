"""Dev script (GPU box): A/B of library builds on the single-plan kernels (5M terrain): CVP whole-grid plan, Dijkstra, inflation.
  python tools/gpu_single2.py <grid side> <lib> [<lib> ...]      lib = path of a libmeshnav_b200.so build ('-' = the in-tree one)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def worker(n, lib):
    sys.path.insert(0, ROOT)
    from mesh_navigation_b200 import _lib
    if lib != "-": _lib.LIB_PATH = os.path.join(ROOT, lib)
    import zlib, numpy as np
    from mesh_navigation_b200 import synth
    from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner, DijkstraMeshPlanner
    pos, faces = synth.grid_mesh(n, n, terrain=True, seed=42)
    mm = MeshMap(pos, faces); mm.setCosts(np.zeros(mm.V, np.float32), mm.edgeDistances())
    c = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
    i, j = min(c % n, n - 2), min(c // n, n - 2); sf = 2 * (j * (n - 1) + i); sp = pos[faces[sf]].mean(0).astype(np.float32)
    pl = CVPMeshPlanner(mm)
    best = 1e9
    for it in range(4):
        g = pl.waveFrontPropagation(sf, sp); best = min(best, g["kernel_ms"])
    d = DijkstraMeshPlanner(mm); bd = 1e9
    for it in range(3):
        gd = d.dijkstra(int(faces[sf][0])); bd = min(bd, gd["kernel_ms"])
    print(f"n={n} lib={lib}: cvp kernel {best:.2f} ms rounds {g['rounds']} recomp/V {g['recomputes']/mm.V:.2f} crc {zlib.crc32(g['dist'].tobytes()):08x} | dijkstra {bd:.2f} ms rounds {gd['rounds']}", flush=True)

if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), sys.argv[3])
    else:
        for lib in sys.argv[2:]:
            r = subprocess.run([sys.executable, __file__, "--worker", sys.argv[1], lib], capture_output=True, text=True, timeout=600)
            print((r.stdout.strip() or ("FAILED: " + r.stderr[-400:])), flush=True)
