"""Dev script (GPU box): large-mesh run (SURVEY 8d config 5 direction): build, fused layers, CVP + Dijkstra full field."""
import sys, time, subprocess
import numpy as np
sys.path.insert(0, '.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner, DijkstraMeshPlanner
n = int(sys.argv[1])
avail_gb = int([l for l in open('/proc/meminfo') if l.startswith('MemAvailable')][0].split()[1]) / 1e6
print(f'host MemAvailable {avail_gb:.0f} GB', flush=True)
if n * n > 30e6 and avail_gb < 150:
    raise SystemExit('not enough host memory for the topology build of this size; skipping')
t0 = time.time(); pos, faces = synth.grid_mesh(n, n, terrain=True); t1 = time.time()
mm = MeshMap(pos, faces); t2 = time.time()
print(f"n={n} V={mm.V} F={mm.F} E={mm.E} mesh synth {t1-t0:.1f}s, mnb_set_mesh (host topology + upload) {t2-t1:.1f}s", flush=True)
ed = mm.edgeDistances(); mm.setCosts(np.zeros(mm.V, np.float32), ed)
print(subprocess.run(["nvidia-smi", "--query-gpu=memory.used,memory.total", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip(), flush=True)
L = mm.computeLayers(); L = mm.computeLayers()
print(f"layers kernel_ms={L['kernel_ms']:.2f} ({837 * mm.V / (L['kernel_ms'] * 1e-3) / 1e9:.0f} GB/s algorithmic)", flush=True)
del L
c = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
sf = int(2 * ((c // n) * (n - 1) + (c % n))); sp = pos[faces[sf]].mean(0).astype(np.float32)
for rep in range(2):
    g = CVPMeshPlanner(mm).waveFrontPropagation(sf, sp)
print(f"cvp full field kernel_ms={g['kernel_ms']:.2f} rounds={g['rounds']} recomp/V={g['recomputes']/mm.V:.2f} settled={g['settled']} "
      f"vertices/s={g['settled']/(g['kernel_ms']*1e-3):.3e} finite={int(np.isfinite(g['dist']).sum())} max={np.nanmax(g['dist'][np.isfinite(g['dist'])]):.2f}", flush=True)
del g
for rep in range(2):
    d = DijkstraMeshPlanner(mm).dijkstra(int(c))
print(f"dijkstra full field kernel_ms={d['kernel_ms']:.2f} rounds={d['rounds']} vertices/s={mm.V/(d['kernel_ms']*1e-3):.3e}", flush=True)
print(subprocess.run(["nvidia-smi", "--query-gpu=memory.used,memory.total", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip(), flush=True)
mm.close()
