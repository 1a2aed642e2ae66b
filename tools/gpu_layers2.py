"""Dev script (GPU box): the k_layers variants on the bench mesh -- kernel ms per mode, bit-equality against mode 0.
  python tools/gpu_layers2.py [grid side, default 2240]"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, '.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2240
pos, faces = synth.grid_mesh(n, n, terrain=True, seed=42)
mm = MeshMap(pos, faces)
mm.L.mnb_debug_set_layers_smem.argtypes = [C.c_void_p, C.c_int32]
base = None
for mode in (0, 2, 5, 6, 7, 8, 9, 2, 5):
    mm.L.mnb_debug_set_layers_smem(mm._ctx, mode)
    best = 1e9
    for it in range(3):
        Ly = mm.computeLayers(); best = min(best, Ly["kernel_ms"])
    if base is None: base = Ly
    same = all((Ly[k].view(np.uint32) == base[k].view(np.uint32)).all() for k in ("height_diff", "roughness", "steepness", "ridge", "combined")) and (Ly["lethal_mask"] == base["lethal_mask"]).all()
    print(f"[layers {n}x{n}] mode={mode}: kernel {best:.3f} ms -> {837 * mm.V / (best * 1e-3) / 1e9:.0f} GB/s algorithmic, identical to mode 0: {same}", flush=True)
