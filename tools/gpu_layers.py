"""Dev script (GPU box): fused layers kernel timing."""
import sys, numpy as np
sys.path.insert(0, '.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap
n = int(sys.argv[1])
pos, faces = synth.grid_mesh(n, n, terrain=True)
mm = MeshMap(pos, faces)
for rep in range(3):
    L = mm.computeLayers()
    print(f"n={n} layers kernel_ms={L['kernel_ms']:.3f} lethal={int((L['lethal_mask'] != 0).sum())} hbm_frac={837 * mm.V / (L['kernel_ms'] * 1e-3) / 1e9 / 6574.1:.4f}", flush=True)
mm.close()
