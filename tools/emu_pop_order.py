"""Dev script (no GPU): compares the POP ORDER implied by the engine's labels (level stacks rebuilt from the label words,
the side arrays and the level pool) with the oracle's recorded pop sequence, on a fuzz fixture, and prints the first
vertices whose order differs.   python tools/emu_pop_order.py <fixture name> [cluster] [norobot]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); os.chdir(ROOT)
os.environ.setdefault("MNB_EMU_SMS", "4")
from tests.emu.run_suite import build
from mesh_navigation_b200 import _lib
_lib.LIB_PATH = build()
import numpy as np
from oracle import oracle as O
from mesh_navigation_b200 import api

name = sys.argv[1]
d = np.load(name if name.endswith(".npz") else f"tests/golden/fuzz_{name}.npz")
pos, faces, vc, w, inv, sf, sp, rf, cl = (d["pos"], d["faces"], d["vc"], d["w"], d["inv"], int(d["sf"]), d["sp"], int(d["rf"]), float(d["cl"]))
if "norobot" in sys.argv: rf = -1
if inv.size == 0: inv = None
om = O.OracleMesh(pos, faces)
pop = np.full(om.V, 0xffffffff, np.uint32)
O._lib.orc_debug_set_pop_buffer(pop.ctypes.data_as(C.c_void_p))
ref = om.cvp(w, vc, sf, sp, rf, invalid=inv, cost_limit=cl)
O._lib.orc_debug_set_pop_buffer(None)
mm = api.MeshMap(pos, faces); mm.setCosts(vc, w, inv)
cluster = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].lstrip('-').isdigit() else -1
mm.set_tuning(0.3 if cluster != -1 else 0.0, cluster, 0)
got = api.CVPMeshPlanner(mm, cost_limit=cl).waveFrontPropagation(sf, sp, rf)
V = om.V
lab = np.zeros((V, 4), np.uint32); root = np.zeros(V, np.uint32); ext = np.zeros(V, np.uint32)
NP = 1 << 20; pool = np.zeros(NP, np.uint32)
L = mm.L
p = lambda a: a.ctypes.data_as(C.c_void_p)
L.mnb_debug_get_labels.argtypes = [C.c_void_p, C.c_void_p]
L.mnb_debug_get_label_sides.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
assert L.mnb_debug_get_labels(mm._ctx, p(lab)) == 0 and L.mnb_debug_get_label_sides(mm._ctx, p(root), p(ext), p(pool), NP) == 0
f = lambda u: float(np.uint32(u).view(np.float32))
def stack(v):
    dbits, a1, z, wv = [int(x) for x in lab[v]]
    a2, a3 = z & 0x7fffffff, wv & 0x7fffffff
    r = int(root[v]) if z >> 31 else v
    e = int(ext[v]) if wv >> 31 else 0
    if a2 == 0: return [(f(a1), r)]
    if a3 == 0: return [(f(a1), r), (f(a2), v)]
    if not (e >> 31): return [(f(a1), r), (f(a2), e), (f(a3), v)]
    o = e & 0x7fffffff; n = int(pool[o])
    s = [(f(a1), r), (f(a2), int(pool[o + 1])), (f(a3), int(pool[o + 2]))]
    for i in range(4, n + 1): s.append((f(pool[o + 3 + 2 * (i - 4)]), int(pool[o + 4 + 2 * (i - 4)])))
    return s
fin = np.where(np.isfinite(got["dist"]))[0]
stacks = {int(v): stack(int(v)) for v in fin}
order = sorted(stacks, key=lambda v: stacks[v])
ref_order = [int(v) for v in np.argsort(pop, kind="stable") if pop[v] != 0xffffffff]
bad = np.where(got["dist"].view(np.uint32) != ref["dist"].view(np.uint32))[0]
print(f"V {V} mismatching potentials {bad.size} deep {got['deep_labels']} labelled gpu {len(order)} oracle {len(ref_order)}")
gi = {v: i for i, v in enumerate(order)}
n_show = 0
for i, v in enumerate(ref_order):
    if i >= len(order) or order[i] != v:
        print(f"first divergence at pop #{i}: oracle pops {v} (d_ref {ref['dist'][v]:.7g}, gpu d {got['dist'][v]:.7g}, gpu stack {stacks.get(v)}, gpu position {gi.get(v)})")
        if i < len(order):
            u = order[i]; print(f"   gpu pops {u} there: d_ref {ref['dist'][u]:.7g} gpu d {got['dist'][u]:.7g} stack {stacks[u]} oracle position {int(pop[u])}")
        for k in range(max(0, i - 3), i): print(f"   before: #{k} {ref_order[k]} stack {stacks[ref_order[k]]} d {ref['dist'][ref_order[k]]:.7g}")
        break
else:
    print("pop order identical")
