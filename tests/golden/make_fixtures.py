"""Regenerates tests/golden/*.npz from the oracle (python -m tests.golden.make_fixtures).

The reference cannot be built or imported in this container (lvr2 / ROS 2 absent), so these are
ORACLE outputs on tiny seeded meshes -- regression fixtures for the restatement, not reference output.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.util import centre_seed, disc_lethals, mesh_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run_case(name):
    terrain = "terrain" in name
    pos, faces = mesh_case(30, terrain)
    m = O.OracleMesh(pos, faces)
    ed = m.edge_distances()
    rng = np.random.default_rng(3)
    vc = (rng.random(m.V) * 0.6).astype(np.float32)
    w = m.edge_weights(vc, ed, 1.0)
    v, f, sp = centre_seed(pos, faces, (0.4, 0.55))
    if name.startswith("cvp"):
        r = m.cvp(w, vc, f, sp)
        return dict(dist=r["dist"], pred=r["pred"], direction=r["direction"], cut=r["cutting_face"].astype(np.int32))
    if name.startswith("dijkstra"):
        r = m.dijkstra(w, vc, v)
        return dict(dist=r["dist"], pred=r["pred"])
    if name.startswith("inflation"):
        le = disc_lethals(pos, 3, 0.25, seed=5)
        r = m.inflation(ed, le)
        return dict(dist=r["dist"], cost=r["cost"])
    if name.startswith("path"):
        # f1/f2 rows: vector map, back-tracking walk and localisation on the same cost-weighted terrain
        rv, rf, rp = centre_seed(pos, faces, (0.85, 0.2))
        r = m.cvp(w, vc, f, sp, rf)
        vn = m.layers()["vertex_normals"]
        vm = m.cvp_vector_map(vn, r["pred"], r["direction"], r["cutting_face"])
        rc, pp, pf = m.cvp_backtrack(vm, sp, f, rp, rf, 0.4)
        q = np.stack([rp, sp, pos[17] + np.float32(0.02), pos[faces[100]].mean(0) + np.float32([0, 0, 0.05])]).astype(np.float32)
        nv, fc, ba = m.locate(q)
        return dict(vector_map=vm, outcome=np.int32(rc), path_pos=pp, path_face=pf, loc_vertex=nv, loc_face=fc, loc_bary=ba,
                    dijkstra_vector_map=m.dijkstra_vector_map(m.dijkstra(w, vc, v)["pred"]))
    if name.startswith("dynamic"):
        # f3/f4 rows: two obstacle configurations in a row -- repulsive field, update set, incremental cost / weight update
        le0 = disc_lethals(pos, 3, 0.25, seed=5); le1 = disc_lethals(pos, 3, 0.25, seed=6)
        inv = (rng.random(m.V) < 0.02).astype(np.uint8)
        i0 = m.inflation(ed, le0, invalid=inv, with_vectors=True)
        i1 = m.inflation(ed, le1, invalid=inv, with_vectors=True)
        upd = O.inflation_update_set(i1["cost"], i0["cost"])
        final = vc.copy()
        O.max_combination_update([vc, i1["cost"]], [0.0, 0.0], [None, None], upd, final, None)
        vc2 = vc.copy(); O.layer_changed(final, 0.0, upd, vc2)
        w2 = w.copy(); m.update_edge_weights(vc2, ed, 1.0, upd, w2)
        fq = np.arange(0, m.F, 7, dtype=np.uint32); b = np.tile(np.float32([0.5, 0.3, 0.2]), (fq.size, 1))
        at = m.inflation_vector_at(fq, b, i1["dist"], i1["vectors"])
        return dict(vectors0=i0["vectors"], vectors1=i1["vectors"], update_set=upd, final=final, vertex_costs=vc2, edge_weights=w2,
                    vector_at=at)
    if name.startswith("raycast"):
        # f3 remainder: ray casting, the obstacle layer's point-cloud update (two clouds in a row) and calcNormalClearance on a
        # terrain with a tilted roof patch above its middle
        rpos, rfaces = mesh_case(12, False, seed=5)
        rpos = rpos.copy(); rpos[:, 0] += 0.9; rpos[:, 1] += 0.9
        rpos[:, 2] = float(pos[:, 2].max()) + 0.6 + 0.1 * (rpos[:, 0] - 0.9)
        P = np.vstack([pos, rpos]).astype(np.float32); Fc = np.vstack([faces, rfaces + pos.shape[0]]).astype(np.uint32)
        mr = O.OracleMesh(P, Fc)
        o = (P.min(0) - 0.3 + rng.random((300, 3)) * (P.max(0) - P.min(0) + 0.6)).astype(np.float32)
        d = rng.normal(size=(300, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        d[:60] = np.float32([0, 0, -1]); d[60:80] = np.float32([0, 0, 1]); o[:20] = P[::37][:20] + np.float32([0, 0, 0.5])
        rc = mr.cast_rays(o, d)
        mask = np.zeros(mr.V, np.uint8)
        T = np.float32([[0.96, -0.28, 0, 1.5], [0.28, 0.96, 0, 1.4], [0, 0, 1, float(P[:, 2].mean()) + 1.0]])
        ax = np.float32([0.06, -0.03, -1.0]); ax = (ax / np.linalg.norm(ax)).astype(np.float32)
        out = {}
        for step in range(2):
            pts = (rng.normal(size=(500, 3)) * np.float32([0.8, 0.8, 0.5])).astype(np.float32)
            le, ch = mr.obstacle_update(pts, T, ax, 2.0, 0.9, mask)
            out[f"lethals{step}"] = le; out[f"changed{step}"] = ch
        vn = mr.layers()["vertex_normals"]
        return dict(hit=rc["hit"].astype(np.uint32), dist=rc["dist"], face=rc["face"], point=rc["point"], clearance=mr.normal_clearance(vn), **out)
    raise KeyError(name)


if __name__ == "__main__":
    for n in ["cvp_planar30", "cvp_terrain30", "dijkstra_terrain30", "inflation_terrain30", "path_terrain30", "dynamic_terrain30", "raycast_terrain30"]:
        np.savez_compressed(os.path.join(HERE, n + ".npz"), **run_case(n))
        print("wrote", n)
