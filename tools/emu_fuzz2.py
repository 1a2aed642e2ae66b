"""Dev script (no GPU): randomised differential test of the NON-wavefront entry points on the interpreted kernels --
fused layers, localisation, vector maps + back-tracking, incremental cost updates vs a fresh install, inflation update set.
  python tools/emu_fuzz2.py [n_cases] [seed] [only]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MNB_EMU_SMS", "4")
from tests.emu.run_suite import build
from mesh_navigation_b200 import _lib
_lib.LIB_PATH = build()
import numpy as np
from oracle import oracle as O
from mesh_navigation_b200 import synth, api
from tests.util import delaunay_mesh, disc_lethals

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
ONLY = int(sys.argv[3]) if len(sys.argv) > 3 else -1
bad = 0
for case in range(N):
    r = np.random.default_rng(rng.integers(1 << 62))
    if ONLY >= 0 and case != ONLY:
        continue
    kind = r.choice(["grid", "planar", "delaunay"])
    if kind == "delaunay":
        pos, faces = delaunay_mesh(int(r.integers(600, 2500)), seed=int(r.integers(1 << 30)), with_hub=bool(r.integers(2)))
    else:
        n = int(r.integers(20, 90))
        pos, faces = synth.grid_mesh(n, n, terrain=kind == "grid", seed=int(r.integers(1 << 30)), z_scale=float(r.choice([2.0, 6.0])))
    om = O.OracleMesh(pos, faces); mm = api.MeshMap(pos, faces); V = om.V
    ed = om.edge_distances()
    msg = []
    # ---- fused layers
    P = api._lib.LayerParams.defaults(); PO = O.LayerParams.defaults()
    if r.integers(2):
        for Q in (P, PO):
            Q.roughness_radius = 0.2; Q.ridge_radius = 0.45; Q.height_diff_threshold = 0.05
    cl = (r.random(V) * 1.2).astype(np.float32) if r.integers(2) else None
    try:
        gl = mm.computeLayers(P, cl); rl = om.layers(PO, cl)
        for k in ("height_diff", "ridge", "border", "clearance"):
            if (gl[k].view(np.uint32) != rl[k].view(np.uint32)).any(): msg.append(f"layer {k} differs at {(gl[k].view(np.uint32) != rl[k].view(np.uint32)).sum()}")
        for k in ("roughness", "steepness"):
            if not np.allclose(gl[k], rl[k], rtol=1e-5, atol=2e-6): msg.append(f"layer {k} max abs {np.abs(gl[k]-rl[k]).max():.2e}")
        if not np.allclose(gl["combined"], rl["combined"], rtol=1e-5, atol=2e-6): msg.append("combined differs")
    except api.MeshNavError as ex:
        msg.append(f"layers: {ex}")
    # ---- localisation
    q = (pos[r.integers(V, size=40)] + r.normal(0, 0.05, (40, 3))).astype(np.float32)
    gv, gf, gb = mm.locate(q); rv, rf_, rb = om.locate(q)
    if (gv != rv).any() or (gf != rf_).any() or (gb.view(np.uint32) != rb.view(np.uint32)).any(): msg.append("locate differs")
    # ---- plan + vector map + back-tracking
    vc = (r.random(V) * 0.6).astype(np.float32) if r.integers(2) else np.zeros(V, np.float32)
    factor = float(r.choice([0.0, 1.0]))
    w = om.edge_weights(vc, ed, factor); mm.setCosts(vc, w)
    sf = int(r.integers(om.F)); sp = pos[faces[sf]].mean(0).astype(np.float32)
    rf = int(r.integers(om.F)); rp = pos[faces[rf]].mean(0).astype(np.float32)
    pl = api.CVPMeshPlanner(mm)
    g = pl.waveFrontPropagation(sf, sp, rf); ref = om.cvp(w, vc, sf, sp, rf)
    if (g["dist"].view(np.uint32) != ref["dist"].view(np.uint32)).any(): msg.append("cvp differs")
    vn = om.layers()["vertex_normals"]
    gvm = pl.computeVectorMap(g["pred"], g["direction"], g["cutting_face"]); rvm = om.cvp_vector_map(vn, ref["pred"], ref["direction"], ref["cutting_face"])
    if not np.array_equal(np.isnan(gvm), np.isnan(rvm)) or not np.allclose(np.nan_to_num(gvm), np.nan_to_num(rvm), atol=5e-6): msg.append("vector map differs")
    if g["outcome"] == 0:
        bt = pl.backtrack(rp, rf); rc, pp, pf = om.cvp_backtrack(gvm, sp, sf, rp, rf)       # the oracle walks the DEVICE's field
        if bt["outcome"] != rc or bt["positions"].shape != pp.shape or (bt["positions"].view(np.uint32) != pp.view(np.uint32)).any() or (bt["faces"] != pf).any():
            msg.append(f"backtrack outcome {bt['outcome']}/{rc} points {len(bt['positions'])}/{len(pp)}")
    # ---- incremental update == fresh install, several rounds
    fresh = api.MeshMap(pos, faces)
    vcur, wcur = vc.copy(), w.copy()
    for it in range(3):
        ch = r.choice(V, int(r.integers(1, max(2, V // 5))), replace=False).astype(np.uint32)
        nv = (r.random(ch.size) * 1.5).astype(np.float32); nv[r.random(ch.size) < 0.05] = np.inf
        f2 = float(r.choice([0.0, factor, 2.0]))
        layer = np.full(V, np.nan, np.float32); layer[ch] = nv
        O.layer_changed(layer, 0.0, ch, vcur); om.update_edge_weights(vcur, ed, f2, ch, wcur)
        mm.layerChanged(ch, nv, f2)
        a, b = mm.costs()
        if (a.view(np.uint32) != vcur.view(np.uint32)).any() or (b.view(np.uint32) != wcur.view(np.uint32)).any(): msg.append(f"update {it}: costs/weights differ")
    fresh.setCosts(vcur, wcur)
    g1 = api.CVPMeshPlanner(mm, cost_limit=2.0).waveFrontPropagation(sf, sp); g2 = api.CVPMeshPlanner(fresh, cost_limit=2.0).waveFrontPropagation(sf, sp)
    if (g1["dist"].view(np.uint32) != g2["dist"].view(np.uint32)).any() or (g1["pred"] != g2["pred"]).any(): msg.append("plan on patched tables != fresh install")
    d1 = api.DijkstraMeshPlanner(mm, cost_limit=2.0).dijkstra(int(faces[sf][0])); d2 = api.DijkstraMeshPlanner(fresh, cost_limit=2.0).dijkstra(int(faces[sf][0]))
    if (d1["dist"].view(np.uint32) != d2["dist"].view(np.uint32)).any() or (d1["pred"] != d2["pred"]).any(): msg.append("dijkstra on patched tables != fresh install")
    # ---- inflation update set over two configurations
    infl = api.InflationLayer(mm); old = None
    for it in range(2):
        le = disc_lethals(pos, int(r.integers(0, 6)), 0.25, seed=int(r.integers(1 << 30))) if True else None
        ri = om.inflation(ed, le); gi = infl.onInputChanged(le)
        upd = O.inflation_update_set(ri["cost"], old); old = ri["cost"]
        if not np.array_equal(gi["changed"], upd): msg.append(f"inflation update set {it} differs")
    # ---- ray casting, obstacle layer, normal clearance (a tilted roof patch is added for a third of the cases)
    rpos, rfaces = pos, faces
    if r.integers(3) == 0:
        m = int(r.integers(6, 14))
        qp, qf = synth.grid_mesh(m, m, terrain=False, seed=int(r.integers(1 << 30)))
        qp = qp.copy(); ctr = pos[:, :2].mean(0)
        qp[:, 0] += ctr[0] - 0.05 * m; qp[:, 1] += ctr[1] - 0.05 * m
        qp[:, 2] = float(pos[:, 2].max()) + float(r.random() * 1.0 + 0.2) + float(r.normal() * 0.2) * (qp[:, 0] - ctr[0])
        rpos = np.vstack([pos, qp]).astype(np.float32); rfaces = np.vstack([faces, qf + pos.shape[0]]).astype(np.uint32)
    omr = O.OracleMesh(rpos, rfaces) if rpos is not pos else om
    mr = api.MeshMap(rpos, rfaces) if rpos is not pos else mm
    nr = 150
    lo, hi = rpos.min(0) - 0.3, rpos.max(0) + 0.3
    ro = (lo + r.random((nr, 3)) * (hi - lo)).astype(np.float32)
    rd = r.normal(size=(nr, 3)); rd = (rd / np.linalg.norm(rd, axis=1, keepdims=True)).astype(np.float32)
    rd[:40] = np.float32([0, 0, -1]); ro[:15] = rpos[r.integers(rpos.shape[0], size=15)] + np.float32([0, 0, 0.6])
    gr, rr = mr.castRays(ro, rd), omr.cast_rays(ro, rd)
    if (gr["hit"] != rr["hit"]).any() or (gr["face"] != rr["face"]).any() or (gr["dist"].view(np.uint32) != rr["dist"].view(np.uint32)).any() \
            or (gr["point"].view(np.uint32) != rr["point"].view(np.uint32)).any():
        msg.append(f"ray cast differs ({int((gr['face'] != rr['face']).sum())} faces)")
    ol = api.ObstacleLayer(mr, robot_height=float(r.choice([0.3, 0.8, 5.0])), max_obstacle_dist=float(r.choice([1.5, 4.0])))
    mask = np.zeros(omr.V, np.uint8)
    for it in range(2):
        npts = int(r.integers(0, 400))
        pts = (r.normal(size=(npts, 3)) * np.float32([1.0, 1.0, 0.5])).astype(np.float32)
        ang = float(r.random() * 6.28)
        T = np.float32([[np.cos(ang), -np.sin(ang), 0, rpos[:, 0].mean()], [np.sin(ang), np.cos(ang), 0, rpos[:, 1].mean()], [0, 0, 1, rpos[:, 2].max() + 0.5]])
        ax = np.float32([r.normal() * 0.05, r.normal() * 0.05, -1.0]); ax = (ax / np.linalg.norm(ax)).astype(np.float32)
        rle, rch = omr.obstacle_update(pts, T, ax, ol.config.max_obstacle_dist, ol.config.robot_height, mask)
        go = ol.processPointCloud(pts, T, ax)
        if not np.array_equal(go["lethals"], rle) or not np.array_equal(go["changed"], rch): msg.append(f"obstacle update {it} differs")
    gvn = mr.vertexNormals()
    if (mr.normalClearance(gvn).view(np.uint32) != omr.normal_clearance(gvn).view(np.uint32)).any(): msg.append("normal clearance differs")
    if mr is not mm: mr.close()
    print(f"case {case}: {kind} V={V}: " + ("ok" if not msg else "MISMATCH " + "; ".join(msg)), flush=True)
    bad += bool(msg)
    mm.close(); fresh.close()
print(f"{N - bad}/{N} cases ok")
sys.exit(1 if bad else 0)
