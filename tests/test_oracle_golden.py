"""CPU tests: pin the oracle against every known answer the reference's own tests hold for the
hot path (mesh_layers/test/inflation_layer_test.cpp) plus derived golden vectors and fixtures."""
import os

import numpy as np
import pytest

from tests.util import centre_seed, mesh_case, disc_lethals

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def triangle(O):
    # inflation_layer_test.cpp:7-23 genTriangle()
    pos = np.array([[0, 0, 0], [0.5, 0, 0], [0, 0.5, 0]], np.float32)
    faces = np.array([[0, 1, 2]], np.uint32)
    return O.OracleMesh(pos, faces)


def test_reference_wave_front_update(oracle_mod):
    """inflation_layer_test.cpp:38-80 test_wave_front_update"""
    O = oracle_mod
    m = triangle(O)
    ew = m.edge_distances()                       # calcEdgeWeights(): Euclidean (:26-36)
    e01 = [i for i, e in enumerate(m.edges.tolist()) if e == [0, 1]][0]
    dist = np.array([0.0, ew[e01], np.inf], np.float32)   # sparse map: v2 has no distance yet
    vec = np.zeros((3, 3), np.float32)
    assert m.inflation_wavefront_update(dist, vec, 5.0, ew, 0, 0, 1, 2) is True      # EXPECT_TRUE :62
    assert dist[2] == np.float32(0.5)                                                # EXPECT_FLOAT_EQ :76
    assert O.fading(dist[2], 0.5, 1.5, 1.0, 0.9, 1.0) == np.float32(0.9)             # :79


def test_reference_fading(oracle_mod):
    """inflation_layer_test.cpp:83-100 test_fading"""
    O = oracle_mod
    f = lambda d: O.fading(d, 0.5, 1.5, 1.0, 0.9, 1.0)
    assert f(0.2) == np.float32(0.9)
    assert 0.0 < f(0.6) < 0.9
    assert abs(f(0.6) - 0.9 * np.exp(-0.1)) < 1e-6
    assert f(2.0) == 0.0
    assert f(0.0) == 1.0          # lethality branch (inflation_layer.cpp:337-338)


def test_derived_cvp_triangle(oracle_mod):
    """SURVEY 8c derived vector: same triangle through CVP waveFrontUpdate -> d[v2]=0.5, pred v0, theta 0"""
    m = triangle(oracle_mod)
    ew = m.edge_distances()
    dist = np.array([0.0, 0.5, np.inf], np.float32)
    pred = np.arange(3, dtype=np.uint32); dr = np.zeros(3, np.float32); cut = -np.ones(3, np.int32)
    assert m.cvp_wavefront_update(ew, 0, 0, 1, 2, dist, pred, dr, cut)
    assert dist[2] == np.float32(0.5) and pred[2] == 0 and dr[2] == 0.0 and cut[2] == 0


def test_edge_weights_formula(oracle_mod):
    """mesh_map.cpp:539-553"""
    pos, faces = mesh_case(12, True)
    m = oracle_mod.OracleMesh(pos, faces)
    ed = m.edge_distances()
    rng = np.random.default_rng(0)
    vc = rng.random(m.V).astype(np.float32)
    vc[5] = np.inf
    w = m.edge_weights(vc, ed, 1.5)
    a, b = m.edges[:, 0], m.edges[:, 1]
    infm = np.isinf(vc[a]) | np.isinf(vc[b])
    assert np.isinf(w[infm]).all() and infm.sum() > 0
    edge_cost = ((ed * (vc[a] + vc[b])).astype(np.float32).astype(np.float64) / 2.0).astype(np.float32)   # :550
    expect = (ed.astype(np.float64) + np.float64(1.5) * edge_cost.astype(np.float64)).astype(np.float32)  # :552
    assert (w[~infm] == expect[~infm]).all()
    assert (m.edge_weights(vc, ed, 0.0)[~infm] == ed[~infm]).all()


def test_planar_sanity(oracle_mod):
    """Dijkstra >= Euclid, CVP ~ Euclid on the planar jittered grid (SURVEY 8c sanity vectors)."""
    pos, faces = mesh_case(100, False)
    m = oracle_mod.OracleMesh(pos, faces)
    ed = m.edge_distances(); vc = np.zeros(m.V, np.float32)
    v, f, sp = centre_seed(pos, faces)
    rd = m.dijkstra(ed, vc, v)
    eu = np.linalg.norm(pos - pos[v], axis=1)
    assert (rd["dist"] >= eu - 1e-4).all() and rd["fixed"] == m.V
    rc = m.cvp(ed, vc, f, sp)
    eu = np.linalg.norm(pos - sp, axis=1)
    far = eu > 1.0
    assert np.max(np.abs(rc["dist"][far] - eu[far]) / eu[far]) < 0.03      # discretisation error of the method
    assert (rc["dist"] >= eu - 1e-5).all()                                  # never shorter than the straight line
    # predecessor trees are consistent
    p = rd["pred"]; nz = p != np.arange(m.V)
    assert (rd["dist"][p[nz]] < rd["dist"][nz]).all()


def test_tie_modes_agree_on_jittered_mesh(oracle_mod):
    """canonical tie-break == plain lvr2-style heap unless two heap keys are bit-identical"""
    pos, faces = mesh_case(60, True)
    m = oracle_mod.OracleMesh(pos, faces)
    ed = m.edge_distances(); vc = np.zeros(m.V, np.float32)
    v, f, sp = centre_seed(pos, faces)
    a = m.cvp(ed, vc, f, sp, canonical_ties=True); b = m.cvp(ed, vc, f, sp, canonical_ties=False)
    assert (a["dist"] == b["dist"]).all()
    a = m.dijkstra(ed, vc, v, canonical_ties=True); b = m.dijkstra(ed, vc, v, canonical_ties=False)
    assert (a["dist"] == b["dist"]).all() and (a["pred"] == b["pred"]).all()


def test_goal_cutoff_and_no_path(oracle_mod):
    pos, faces = mesh_case(50, False)
    m = oracle_mod.OracleMesh(pos, faces)
    ed = m.edge_distances(); vc = np.zeros(m.V, np.float32)
    v, f, sp = centre_seed(pos, faces, (0.25, 0.25))
    rv, rf, _ = centre_seed(pos, faces, (0.6, 0.6))
    full = m.dijkstra(ed, vc, v)
    cut = m.dijkstra(ed, vc, v, robot_vertex=rv)
    assert cut["outcome"] == 0 and cut["expanded"] < full["expanded"]
    lim = np.float32(full["dist"][rv] + 0.3)
    inside = full["dist"] <= lim
    assert (cut["dist"][inside] == full["dist"][inside]).all()
    # wall of lethal cost between seed and robot -> NO_PATH_FOUND (54)
    vc2 = vc.copy(); vc2[(pos[:, 0] > 2.2) & (pos[:, 0] < 2.5)] = 2.0
    assert m.dijkstra(ed, vc2, v, robot_vertex=rv)["outcome"] == 54
    assert m.cvp(ed, vc2, f, sp, robot_face=rf)["outcome"] == 54
    assert m.cvp(ed, vc, f, sp, robot_face=rf)["outcome"] == 0


@pytest.mark.parametrize("name", ["cvp_planar30", "cvp_terrain30", "dijkstra_terrain30", "inflation_terrain30", "path_terrain30",
                                  "dynamic_terrain30", "raycast_terrain30"])
def test_committed_fixtures(oracle_mod, name):
    """Fixtures under tests/golden were generated by tests/golden/make_fixtures.py from this oracle
    (the reference cannot run here); they guard the oracle against silent changes."""
    from tests.golden.make_fixtures import run_case
    path = os.path.join(GOLD, name + ".npz")
    assert os.path.exists(path), "run python -m tests.golden.make_fixtures"
    gold = np.load(path)
    now = run_case(name)
    for k in gold.files:
        a, b = gold[k], now[k]
        assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all(), f"{name}:{k} drifted"


def test_next_rows_sanity(oracle_mod):
    """f1/f2 restatements (cvp:204-239, 920-951; mesh_map.cpp:999-1174; util.cpp:313-347): derived properties on a planar
    mesh -- parity for these rows is unpinned by the reference (no test there), so these are sanity bounds."""
    from tests.util import centre_seed, mesh_case
    pos, faces = mesh_case(60, False)
    m = oracle_mod.OracleMesh(pos, faces)
    ed = m.edge_distances(); vc = np.zeros(m.V, np.float32)
    sv, sf, sp = centre_seed(pos, faces, (0.2, 0.3))
    rv, rf, rp = centre_seed(pos, faces, (0.8, 0.75))
    r = m.cvp(ed, vc, sf, sp, rf)
    assert r["outcome"] == 0
    vn = m.layers()["vertex_normals"]
    vm = m.cvp_vector_map(vn, r["pred"], r["direction"], r["cutting_face"])
    has = ~np.isnan(vm).any(1)
    assert np.allclose(np.linalg.norm(vm[has], axis=1), 1.0, atol=1e-5)
    # on a plane the field points straight at the seed (continuous vector field: that is the point of the CVP)
    to_seed = sp[None, :2] - pos[has, :2]
    far = np.linalg.norm(to_seed, axis=1) > 0.5
    cosang = (vm[has][far, :2] * to_seed[far]).sum(1) / np.linalg.norm(to_seed[far], axis=1)
    assert np.median(cosang) > 0.999 and np.percentile(cosang, 5) > 0.98 and cosang.min() > 0.5   # edge fallbacks at the rim
    # Dijkstra's field only knows edge directions
    dj = m.dijkstra(ed, vc, sv)
    dvm = m.dijkstra_vector_map(dj["pred"])
    assert np.isnan(dvm[sv]).all() and np.allclose(np.linalg.norm(dvm[~np.isnan(dvm).any(1)], axis=1), 1.0, atol=1e-5)
    # back-tracking: robot -> seed, steps of step_width, nearly the straight line on a plane
    rc, pp, pf = m.cvp_backtrack(vm, sp, sf, rp, rf, 0.4)
    assert rc == 0 and (pp[0] == rp).all() and (pp[-1] == sp).all() and pf[0] == rf and pf[-1] == sf
    seg = np.linalg.norm(np.diff(pp, axis=0), axis=1)
    assert np.allclose(seg[:-1], 0.4, atol=2e-3) and seg[-1] <= np.sqrt(0.4) + 1e-3      # cvp:925 compares distance^2 with step_width
    chord = np.linalg.norm(rp - sp)
    assert chord <= seg.sum() <= 1.02 * chord + 0.4
    # localisation: a point inside a face is found in that face with its barycentric coordinates
    f = 1234; b = np.float32([0.2, 0.5, 0.3])
    q = (pos[faces[f]] * b[:, None]).sum(0)
    nv, fc, ba = m.locate(np.stack([q, pos[77], q + np.float32([100, 0, 0])]))
    assert fc[0] == f and np.allclose(ba[0], b, atol=1e-4) and nv[1] == 77 and fc[2] == -1
    assert nv[0] in faces[f]


def test_incremental_update_restatements(oracle_mod):
    """f3 restatements on hand-checkable inputs: MeshMap::layerChanged + updateEdgeWeights (mesh_map.cpp:455-492, 563-618),
    MaxCombinationLayer::onInputChanged (combination_layer.cpp:87-147), the update set of InflationLayer::onInputChanged
    (inflation_layer.cpp:154-164).  Unpinned by the reference (it has no test for them): derived values."""
    O = oracle_mod
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)         # unit square, two triangles
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    m = O.OracleMesh(pos, faces)
    edges = m.edges.tolist()
    e = {tuple(x): i for i, x in enumerate(edges)}
    assert sorted(e) == [(0, 1), (0, 2), (0, 3), (1, 2), (2, 3)]
    ed = m.edge_distances()
    assert ed[e[(0, 2)]] == np.float32(np.sqrt(np.float32(2.0)))
    vc = np.float32([0.0, 0.5, 0.25, 1.0])
    w = m.edge_weights(vc, ed, 2.0)
    # computeEdgeWeights: w = d + f * (d * (c1 + c2) / 2)
    assert w[e[(0, 1)]] == np.float32(1.0 + 2.0 * (1.0 * 0.5 / 2.0))
    assert w[e[(2, 3)]] == np.float32(1.0 + 2.0 * (1.0 * 1.25 / 2.0))
    # layerChanged: vertex 1 gets a new cost, vertex 3 loses its entry (-> default)
    layer = np.float32([np.nan, 0.75, np.nan, np.nan])
    vc2 = vc.copy(); O.layer_changed(layer, 0.125, np.uint32([1, 3]), vc2)
    assert vc2.tolist() == [0.0, 0.75, 0.25, 0.125]
    w2 = w.copy(); m.update_edge_weights(vc2, ed, 2.0, np.uint32([1, 3]), w2)
    assert w2[e[(0, 1)]] == np.float32(1.0 + 2.0 * (0.75 / 2.0)) and w2[e[(1, 2)]] == np.float32(1.0 + 2.0 * (1.0 / 2.0))
    assert w2[e[(0, 3)]] == np.float32(1.0 + 2.0 * (0.125 / 2.0)) and w2[e[(2, 3)]] == np.float32(1.0 + 2.0 * (0.375 / 2.0))
    assert w2[e[(0, 2)]] == w[e[(0, 2)]]                                                # not incident to a changed vertex
    assert (w2.view(np.uint32) == m.edge_weights(vc2, ed, 2.0).view(np.uint32)).all()  # == full recompute
    # +inf endpoint -> +inf weight (:598); a zero factor returns before touching anything (:568-572), unlike computeEdgeWeights
    vc3 = vc2.copy(); vc3[1] = np.inf
    w3 = w2.copy(); m.update_edge_weights(vc3, ed, 2.0, np.uint32([1]), w3)
    assert np.isinf(w3[e[(0, 1)]]) and np.isinf(w3[e[(1, 2)]]) and np.isfinite(w3[e[(0, 2)]])
    w4 = w2.copy(); m.update_edge_weights(vc3, ed, 0.0, np.uint32([1]), w4)
    assert (w4 == w2).all() and np.isinf(m.edge_weights(vc3, ed, 0.0)[e[(0, 1)]])
    # MaxCombination: max over layers of (value or default), starting from 0; lethal = any
    a = np.float32([0.2, -1.0, np.nan, 0.4]); b = np.float32([np.nan, np.nan, np.nan, 0.9])
    costs = np.float32([9, 9, 9, 9]); leth = np.uint8([5, 5, 5, 5])
    O.max_combination_update([a, b], [0.0, 0.3], [np.uint8([0, 0, 1, 0]), None], np.uint32([0, 1, 2]), costs, leth)
    assert costs.tolist() == [np.float32(0.3), np.float32(0.3), np.float32(0.3), 9.0] and leth.tolist() == [0, 0, 1, 5]
    O.max_combination_update([a, b], [0.0, 0.3], [None, None], np.uint32([3]), costs, leth)
    assert costs[3] == np.float32(0.9) and leth[3] == 0
    # AvgCombination (combination_layer.cpp:264-271): weighted sum in layer order, float
    avg = np.float32([9, 9, 9, 9]); al = np.uint8([5, 5, 5, 5])
    O.avg_combination_update([a, b], [0.0, 0.3], [0.5, 2.0], [np.uint8([0, 0, 1, 0]), None], np.uint32([0, 2, 3]), avg, al)
    assert avg[0] == np.float32(0.5) * np.float32(0.2) + np.float32(2.0) * np.float32(0.3)
    assert avg[2] == np.float32(0.5) * np.float32(0.0) + np.float32(2.0) * np.float32(0.3) and avg[1] == 9.0
    assert avg[3] == np.float32(0.5) * np.float32(0.4) + np.float32(2.0) * np.float32(0.9) and al.tolist() == [0, 5, 1, 0]
    # update set: keys(new) U keys(old)
    new = np.float32([np.nan, 1.0, np.nan, 0.5]); old = np.float32([0.1, np.nan, np.nan, 0.2])
    assert O.inflation_update_set(new, old).tolist() == [0, 1, 3] and O.inflation_update_set(new).tolist() == [1, 3]


def test_repulsive_field_restatement(oracle_mod):
    """f4 restatement (inflation_layer.cpp:277-308, 493-521) on the reference test's triangle: v0, v1 lethal, v2 free.
    The face is visited from both lethal vertices through both of their edges: 4 accumulations of the same direction, each
    followed by a normalisation -> every vertex ends with dir = normalize((p2-p1) + (p2-p0))."""
    O = oracle_mod
    m = triangle(O)
    ed = m.edge_distances()
    r = m.inflation(ed, np.uint32([0, 1]), inflation_radius=1.5, inscribed_radius=0.7, with_vectors=True)
    # Sethian update with both sources at 0 and the literal sin-theta (not sin^2) of inflation_layer.cpp:195:
    # a = |v1v2| = sqrt(0.5), b = 0.5, cos = sin = sqrt(0.5): f2 = 0.25, f1 = 0, f0 = -0.25 * 0.5 * sqrt(0.5) -> t = sqrt(-f0 * f2) / f2
    t = np.sqrt(0.25 * 0.5 * np.sqrt(0.5) * 0.25) / 0.25
    assert abs(float(r["dist"][2]) - t) < 1e-6 and abs(t - 0.5946035575) < 1e-9
    d = np.float32([-0.5, 1.0, 0.0]); d = d / np.sqrt((d * d).sum(dtype=np.float32))
    for v in range(3):
        assert np.allclose(r["vectors"][v], d, atol=1e-6), v
    # vectorAt at the three vertices (barycentric unit vectors): lethal -> lethal_value, 0 < d <= inscribed -> inscribed_value
    at = m.inflation_vector_at(np.uint32([0, 0, 0]), np.eye(3, dtype=np.float32), r["dist"], r["vectors"], 0.7, 1.5, 1.0, 0.9)
    assert np.allclose(at[0], d * 1.0, atol=1e-6) and np.allclose(at[2], d * 0.9, atol=1e-6)
    # between the radii: inscribed_value * (cos(alpha) + 1) / 2 with alpha = (sqrt(d) - r_in) / (r_infl - r_in) * pi  (sqrt as written, :508)
    at2 = m.inflation_vector_at(np.uint32([0]), np.float32([[0, 0, 1]]), r["dist"], r["vectors"], 0.25, 1.5, 1.0, 0.9)
    alpha = np.float32((np.sqrt(r["dist"][2]) - 0.25) / (1.5 - 0.25) * np.pi)
    assert np.allclose(at2[0], d * np.float32(0.9) * (np.cos(alpha) + 1) / 2, atol=1e-6)
    # beyond the inflation radius / a vertex without an entry: zero
    far = m.inflation_vector_at(np.uint32([0]), np.float32([[0, 0, 1]]), r["dist"], r["vectors"], 0.1, 0.3, 1.0, 0.9)
    assert (far == 0).all()
    rn = m.inflation(ed, np.uint32([0]), inflation_radius=0.1, with_vectors=True)     # v2 is never reached
    assert np.isinf(rn["dist"][2])
    assert (m.inflation_vector_at(np.uint32([0]), np.float32([[0.3, 0.3, 0.4]]), rn["dist"], rn["vectors"]) == 0).all()


def test_dijkstra_against_an_independent_implementation(oracle_mod):
    """the oracle's DijkstraMeshPlanner restatement against scipy.sparse.csgraph.dijkstra (float64, independent code) on
    config 1 (10k planar mesh) and a cost-weighted terrain: same shortest-path distances up to float32 accumulation, and
    the oracle's predecessor tree reproduces its own distances edge by edge"""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import dijkstra as sp_dijkstra
    rng = np.random.default_rng(2)
    for n, terrain, factor in ((100, False, 0.0), (80, True, 1.5)):
        pos, faces = mesh_case(n, terrain)
        m = oracle_mod.OracleMesh(pos, faces)
        ed = m.edge_distances()
        vc = (rng.random(m.V) * 0.8).astype(np.float32) if factor else np.zeros(m.V, np.float32)
        w = m.edge_weights(vc, ed, factor)
        v, f, sp = centre_seed(pos, faces, (0.25, 0.25))
        r = m.dijkstra(w, vc, v, cost_limit=10.0)
        e = m.edges.astype(np.int64)
        g = coo_matrix((np.r_[w, w].astype(np.float64), (np.r_[e[:, 0], e[:, 1]], np.r_[e[:, 1], e[:, 0]])), shape=(m.V, m.V)).tocsr()
        ref = sp_dijkstra(g, directed=False, indices=v)
        assert np.isfinite(r["dist"]).all()
        assert np.allclose(r["dist"], ref, rtol=2e-6, atol=1e-6)
        # the predecessor tree is consistent: d[v] == fl(d[pred] + w(pred, v)) in float32, root = the seed
        ekey = {(int(a), int(b)): i for i, (a, b) in enumerate(m.edges.tolist())}
        p = r["pred"]
        assert p[v] == v
        idx = np.where(np.arange(m.V) != v)[0][::7]
        for x in idx:
            a, b = (int(p[x]), int(x)) if p[x] < x else (int(x), int(p[x]))
            assert r["dist"][x] == np.float32(r["dist"][p[x]] + w[ekey[(a, b)]])


def test_ray_cast_known_answers(oracle_mod):
    """hand-computed values for the ray-casting restatement (parity unpinned: lvr2 / Embree are not vendored): a unit right
    triangle in the plane z = 0 and a parallel copy at z = 2"""
    O = oracle_mod
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 2], [1, 0, 2], [0, 1, 2]], np.float32)
    faces = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    m = O.OracleMesh(pos, faces)
    o = np.array([[0.25, 0.25, 1.0], [0.25, 0.25, 1.0], [0.25, 0.25, 1.0], [0.5, 0.5, 3.0], [0.75, 0.75, 1.0], [0.25, 0.25, 0.0], [0.25, 0.25, -1.0]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, 1], [1, 0, 0], [0, 0, -1], [0, 0, -1], [0, 0, -1], [0, 0, 1]], np.float32)
    r = m.cast_rays(o, d)
    assert r["hit"].tolist() == [1, 1, 0, 1, 0, 1, 1]
    assert r["face"].tolist()[:2] == [0, 1] and r["dist"][0] == 1.0 and r["dist"][1] == 1.0     # two-sided: hit from above and from below
    assert np.isinf(r["dist"][2]) and r["face"][2] == 0xffffffff and np.isnan(r["point"][2]).all()   # parallel to both planes
    assert r["face"][3] == 1 and r["dist"][3] == 1.0                                      # on the hypotenuse of the upper triangle (u + v == 1): the nearest face wins
    assert r["dist"][5] == 0.0 and r["face"][5] == 0                                       # a point on the surface hits at t = 0
    assert r["face"][6] == 0 and r["dist"][6] == 1.0 and np.allclose(r["point"][6], [0.25, 0.25, 0.0])
    # calcNormalClearance: the lower triangle sees the upper one 2 m above along +z; the upper one sees nothing
    vn = np.tile(np.float32([0, 0, 1]), (6, 1))
    cl = m.normal_clearance(vn)
    assert cl[:3].tolist() == [2.0, 2.0, 2.0] and np.isinf(cl[3:]).all()
    # obstacle layer: points 0.4 m above the lower triangle, robot height 0.5 -> its three vertices; 0.3 -> too far
    mask = np.zeros(6, np.uint8)
    T = np.hstack([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)])
    le, ch = m.obstacle_update(np.float32([[0.2, 0.2, 0.4], [5, 5, 0.4]]), T, np.float32([0, 0, -1]), 10.0, 0.5, mask)
    assert le.tolist() == [0, 1, 2] and ch.tolist() == [0, 1, 2]
    le, ch = m.obstacle_update(np.float32([[0.2, 0.2, 0.4]]), T, np.float32([0, 0, -1]), 10.0, 0.3, mask)
    assert le.size == 0 and ch.tolist() == [0, 1, 2]
    le, ch = m.obstacle_update(np.float32([[0.2, 0.2, 0.4]]), T, np.float32([0, 0, -1]), 0.3, 0.5, mask)    # beyond max_obstacle_dist
    assert le.size == 0 and ch.size == 0


@pytest.mark.parametrize("case", ["geometric", "cost_weighted", "walls_and_cutoff"])
def test_cvp_against_an_independent_restatement(oracle_mod, case):
    """the C++ oracle against tests/cvp_reference_py.py -- a second restatement of CVPMeshPlanner::waveFrontPropagation written
    in plain Python straight from the reference source: potentials, predecessors, directions and cutting faces must agree to
    the last bit (both evaluate the update in double with libm's sqrt / acos and store floats)"""
    from tests.cvp_reference_py import wave_front_propagation
    O = oracle_mod
    pos, faces = mesh_case(26, True)
    m = O.OracleMesh(pos, faces)
    ed = m.edge_distances()
    rng = np.random.default_rng({"geometric": 1, "cost_weighted": 2, "walls_and_cutoff": 3}[case])
    vc = np.zeros(m.V, np.float32); w = ed; inv = None; robot = -1; cl = 1.0
    v, f, sp = centre_seed(pos, faces, (0.35, 0.6))
    if case != "geometric":
        vc = (rng.random(m.V) * 0.7).astype(np.float32)
        w = m.edge_weights(vc, ed, 2.5)                          # non-causal updates: back-steps occur
    if case == "walls_and_cutoff":
        vc[rng.random(m.V) < 0.12] = 1.5                         # over the cost limit
        inv = (rng.random(m.V) < 0.03).astype(np.uint8)
        for x in faces[f]:
            vc[x] = 0.1; inv[x] = 0
        w = m.edge_weights(vc, ed, 1.0)
        rv, robot, rp = centre_seed(pos, faces, (0.8, 0.25))
    ref = m.cvp(w, vc, f, sp, robot_face=robot, cost_limit=cl, invalid=inv)
    got = wave_front_propagation(pos, faces, m.edges, w, vc, f, sp, robot_face=robot, invalid=inv, cost_limit=cl)
    assert (got["dist"].view(np.uint32) == ref["dist"].view(np.uint32)).all()
    assert (got["pred"] == ref["pred"]).all()
    assert (got["cutting_face"] == ref["cutting_face"].astype(np.int64)).all()
    assert (got["direction"].view(np.uint32) == ref["direction"].view(np.uint32)).all()
    if case != "geometric":
        assert ref.get("backsteps", 1) > 0


@pytest.mark.parametrize("case", ["discs", "discs_and_invalid", "wide_radius"])
def test_inflation_against_an_independent_restatement(oracle_mod, case):
    """the C++ oracle against tests/inflation_reference_py.py -- a second restatement of InflationLayer::waveCostInflation
    (Sethian update, fallback, fading, the heap loop with never-fixed invalid vertices) in plain Python from the reference
    source: distances and riskiness values must agree to the last bit"""
    from tests.inflation_reference_py import wave_cost_inflation
    O = oracle_mod
    pos, faces = mesh_case(28, True)
    m = O.OracleMesh(pos, faces)
    ed = m.edge_distances()
    rng = np.random.default_rng(4)
    le = disc_lethals(pos, 3, 0.22, seed=9)
    inv = (rng.random(m.V) < 0.05).astype(np.uint8) if case == "discs_and_invalid" else None
    kw = dict(inscribed_radius=0.3, inflation_radius=0.9, cost_scaling_factor=2.0) if case == "wide_radius" else {}
    ref = m.inflation(ed, le, invalid=inv, **kw)
    got = wave_cost_inflation(pos, faces, m.edges, ed, le, invalid=inv, **kw)
    assert (got["dist"].view(np.uint32) == ref["dist"].view(np.uint32)).all()
    both_nan = np.isnan(got["cost"]) & np.isnan(ref["cost"])
    assert ((got["cost"].view(np.uint32) == ref["cost"].view(np.uint32)) | both_nan).all()
    assert np.isfinite(ref["dist"]).sum() > le.size + 50
