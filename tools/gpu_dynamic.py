"""Dev script (GPU box): one dynamic-obstacle cycle on the 5M map with every array resident on the device (torch tensors,
MNB_PTR_DEVICE): InflationLayer::onInputChanged -> MaxCombinationLayer::onInputChanged -> MeshMap::layerChanged
(inflation_layer.cpp:97-179, combination_layer.cpp:87-147, mesh_map.cpp:454-493,563-618), wall times around
torch.cuda.synchronize(), and the same state obtained by a full re-install for comparison.
  python tools/gpu_dynamic.py [grid side]"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from mesh_navigation_b200 import synth, _lib
from mesh_navigation_b200.api import MeshMap, InflationLayer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2240
pos, faces = synth.grid_mesh(n, n, terrain=True, seed=42)
mm = MeshMap(pos, faces); V = mm.V; L = mm.L; ctx = mm._ctx
ed = mm.edgeDistances(); mm.setCosts(np.zeros(V, np.float32), ed)
Ly = mm.computeLayers()
static_h = Ly["combined"]; base = np.where(Ly["lethal_mask"] != 0)[0]
nd = max(2, int(round(1000 * (n / 2240.0) ** 2)))
le0 = np.union1d(base, synth.disc_lethals_grid(pos, n, n, nd, 0.3, seed=7)).astype(np.uint32)
le1 = np.union1d(base, synth.disc_lethals_grid(pos, n, n, nd, 0.3, seed=8)).astype(np.uint32)
infl = InflationLayer(mm)
r0 = infl.onInputChanged(le0)
final0 = np.maximum(static_h, np.nan_to_num(r0["cost"], nan=0.0)).astype(np.float32)
mm.computeEdgeWeights(final0, 1.0, want_output=False)
dev = torch.device("cuda")
d_static = torch.from_numpy(static_h).to(dev); d_final = torch.from_numpy(final0).to(dev)
d_le1 = torch.from_numpy(le1.astype(np.int64)).to(dev).to(torch.int32)      # (uint32 ids as int32 bit patterns)
d_dist = torch.empty(V, dtype=torch.float32, device=dev); d_cost = torch.empty(V, dtype=torch.float32, device=dev)
d_changed = torch.empty(V, dtype=torch.int32, device=dev); d_vec = torch.empty((V, 3), dtype=torch.float32, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
nch = C.c_uint32(0)
defaults = np.zeros(2, np.float32)
mm.use_device_pointers(True)
def cycle(le_t, n_le, with_field):
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    assert L.mnb_inflation_update(ctx, p(le_t), n_le, None, C.byref(infl.config), p(d_dist), p(d_cost), p(d_changed), C.byref(nch)) == 0
    torch.cuda.synchronize(); t1 = time.perf_counter(); t["inflation_update"] = 1e3 * (t1 - t0); t["inflation_kernel"] = mm.stats()["kernel_ms"]
    lc = (C.c_void_p * 2)(d_static.data_ptr(), d_cost.data_ptr())
    assert L.mnb_max_combination_update(ctx, 2, lc, defaults.ctypes.data_as(C.c_void_p), None, nch.value, p(d_changed), p(d_final), None) == 0
    torch.cuda.synchronize(); t2 = time.perf_counter(); t["max_combination"] = 1e3 * (t2 - t1)
    assert L.mnb_update_vertex_costs(ctx, nch.value, p(d_changed), p(d_final), 1, 0.0, 1.0) == 0
    torch.cuda.synchronize(); t3 = time.perf_counter(); t["layer_changed"] = 1e3 * (t3 - t2)
    if with_field:
        assert L.mnb_inflation_vector_map(ctx, p(d_vec)) == 0
        torch.cuda.synchronize(); t["vector_field"] = 1e3 * (time.perf_counter() - t3)
    t["total_without_field"] = 1e3 * (t3 - t0)
    return t
d_le0 = torch.from_numpy(le0.astype(np.int64)).to(dev).to(torch.int32)
for rep in range(3):
    ta = cycle(d_le1, le1.size, rep == 2); tb = cycle(d_le0, le0.size, False)
print(f"[dynamic {n}x{n}] changed {nch.value} lethals {le1.size}: " + ", ".join(f"{k} {v:.3f} ms" for k, v in ta.items()), flush=True)
print(f"[dynamic {n}x{n}] back:    " + ", ".join(f"{k} {v:.3f} ms" for k, v in tb.items()), flush=True)
# correctness: the incrementally maintained tables equal a full re-install of the same final costs
ta = cycle(d_le1, le1.size, False)
mm.use_device_pointers(False)
gvc, gw = mm.costs()
final1 = d_final.cpu().numpy()
full_w = mm.computeEdgeWeights(final1, 1.0)
ref1 = np.maximum(static_h, np.nan_to_num(infl.onInputChanged(le1)["cost"], nan=0.0)).astype(np.float32)
print("incremental == full re-install:", bool((gw.view(np.uint32) == full_w.view(np.uint32)).all() and (gvc.view(np.uint32) == final1.view(np.uint32)).all()),
      "| combined costs == host Max combination:", bool((final1.view(np.uint32) == ref1.view(np.uint32)).all()))
