"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI
(libmeshnav_b200.so) and is compared with the CPU oracle on the same seeded inputs.

Bars (north_star): Dijkstra distances bit-identical and predecessors bit-exact;
CVP potentials within 1e-4 relative (we additionally report/expect bit-equality,
which the engine's event-ordered replay achieves on these meshes)."""
import numpy as np
import pytest

from tests.util import centre_seed, disc_lethals, face_of_vertex, mesh_case, rel_err

pytestmark = pytest.mark.gpu

CVP_RTOL = 1e-4   # north_star: "within 1e-4 rel (CVP float)"


@pytest.fixture(scope="module")
def api():
    from mesh_navigation_b200 import api as A
    return A


def setup(api, oracle_mod, n, terrain, seed=42, costs=None, factor=0.0, invalid=None):
    pos, faces = mesh_case(n, terrain, seed)
    om = oracle_mod.OracleMesh(pos, faces)
    mm = api.MeshMap(pos, faces)
    assert (mm.edges() == om.edges).all()
    ed = om.edge_distances()
    vc = np.zeros(om.V, np.float32) if costs is None else costs(pos).astype(np.float32)
    w = om.edge_weights(vc, ed, factor)
    gw = mm.computeEdgeWeights(vc, factor)
    assert (gw.view(np.uint32) == w.view(np.uint32)).all(), "computeEdgeWeights differs from mesh_map.cpp:539-553"
    mm.setCosts(vc, w, invalid)
    return pos, faces, om, mm, ed, vc, w


def test_edge_distances_bit_exact(api, oracle_mod):
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, 80, True)
    assert (mm.edgeDistances().view(np.uint32) == ed.view(np.uint32)).all()
    mm.close()


@pytest.mark.parametrize("n,terrain,cluster", [(100, False, 1), (100, False, 8), (64, True, 2), (200, True, 16), (150, True, 4)])
def test_dijkstra_bit_exact(api, oracle_mod, n, terrain, cluster):
    """config 1: DijkstraMeshPlanner single goal, 10k planar mesh (+ terrain variants)"""
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, n, terrain)
    mm.set_tuning(0.3, cluster, 0)
    v, f, sp = centre_seed(pos, faces, (0.25, 0.25))
    ref = om.dijkstra(w, vc, v)
    got = api.DijkstraMeshPlanner(mm).dijkstra(v)
    assert got["outcome"] == ref["outcome"] == 0
    assert (got["dist"].view(np.uint32) == ref["dist"].view(np.uint32)).all()
    assert (got["pred"] == ref["pred"]).all()
    mm.close()


def test_dijkstra_goal_cutoff_costs_invalid(api, oracle_mod):
    """goal_dist cutoff (:293-300), cost_limit (:302), invalid (:328), inf edge weights, NO_PATH (:358)"""
    rng = np.random.default_rng(1)
    costs = lambda pos: np.where(rng.random(pos.shape[0]) < 0.05, 1.5, rng.random(pos.shape[0]) * 0.8)
    pos, faces = mesh_case(120, True)
    invalid = (rng.random(pos.shape[0]) < 0.01).astype(np.uint8)
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, 120, True, costs=costs, factor=1.0, invalid=invalid)
    v, f, sp = centre_seed(pos, faces, (0.3, 0.3))
    invalid[v] = 0
    rv, _, _ = centre_seed(pos, faces, (0.6, 0.7))
    for robot in (-1, rv):
        ref = om.dijkstra(w, vc, v, robot_vertex=robot, invalid=invalid)
        got = api.DijkstraMeshPlanner(mm).dijkstra(v, robot)
        assert got["outcome"] == ref["outcome"]
        assert (got["dist"].view(np.uint32) == ref["dist"].view(np.uint32)).all()
        assert (got["pred"] == ref["pred"]).all()
    # infinite vertex costs -> +inf edge weights (mesh_map.cpp:545) and a wall -> NO_PATH_FOUND
    vc2 = vc.copy(); vc2[(pos[:, 0] > 5.0) & (pos[:, 0] < 5.4)] = np.inf
    w2 = om.edge_weights(vc2, ed, 1.0)
    mm.setCosts(vc2, w2, invalid)
    ref = om.dijkstra(w2, vc2, v, robot_vertex=rv, invalid=invalid)
    got = api.DijkstraMeshPlanner(mm).dijkstra(v, rv)
    assert ref["outcome"] == 54 and got["outcome"] == 54
    assert (got["dist"].view(np.uint32) == ref["dist"].view(np.uint32)).all() and (got["pred"] == ref["pred"]).all()
    mm.close()


def check_cvp(got, ref, exact_aux=True):
    r = rel_err(got["dist"], ref["dist"])
    assert r.max() <= CVP_RTOL, f"max rel err {r.max():.3e}"
    neq = int((got["dist"].view(np.uint32) != ref["dist"].view(np.uint32)).sum())
    if exact_aux and neq == 0:
        assert (got["pred"] == ref["pred"]).all()
        assert (got["cutting_face"] == ref["cutting_face"]).all()
        assert np.abs(got["direction"] - ref["direction"]).max() <= 1e-5
    return neq


@pytest.mark.parametrize("n,terrain,cluster,delta", [(100, False, -1, 0.3), (160, True, -1, 0.3), (100, False, 1, 0.3), (100, False, 8, 0.3), (100, True, 8, 0.1),
                                                     (128, True, 16, 0.6), (200, True, 2, 0.3), (90, False, 4, 1.0)])
def test_cvp_full_field(api, oracle_mod, n, terrain, cluster, delta):
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, n, terrain)
    mm.set_tuning(delta, cluster, 0)
    v, f, sp = centre_seed(pos, faces)
    ref = om.cvp(w, vc, f, sp)
    got = api.CVPMeshPlanner(mm).waveFrontPropagation(f, sp)
    assert got["outcome"] == ref["outcome"] == 0
    assert check_cvp(got, ref) == 0, "potentials are expected to be bit-identical on these meshes"
    assert got["settled"] <= om.V
    mm.close()


def test_cvp_seed_near_vertex_and_border(api, oracle_mod):
    """seed point almost on a vertex (very unequal seed distances) and in a corner face of the mesh"""
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, 90, True)
    for f in (face_of_vertex(faces, 45 * 90 + 45), 0, faces.shape[0] - 1):
        tri = pos[faces[f]]
        sp = (0.98 * tri[0] + 0.01 * tri[1] + 0.01 * tri[2]).astype(np.float32)
        ref = om.cvp(w, vc, f, sp)
        got = api.CVPMeshPlanner(mm).waveFrontPropagation(f, sp)
        check_cvp(got, ref)
    mm.close()


def test_cvp_costs_cutoff_invalid(api, oracle_mod):
    """cost_limit (cvp:757,802), edge_cost_factor weights, invalid (:760,:785), goal cutoff (:754,:763-771), outcome codes"""
    rng = np.random.default_rng(5)
    costs = lambda pos: np.where(rng.random(pos.shape[0]) < 0.04, 1.2, rng.random(pos.shape[0]) * 0.7)
    pos, faces = mesh_case(140, True)
    invalid = (rng.random(pos.shape[0]) < 0.005).astype(np.uint8)
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, 140, True, costs=costs, factor=1.0, invalid=invalid)
    v, f, sp = centre_seed(pos, faces, (0.3, 0.35))
    for x in faces[f]:
        invalid[x] = 0; vc[x] = 0.1
    w = om.edge_weights(vc, ed, 1.0)
    mm.setCosts(vc, w, invalid)
    rv, rf, _ = centre_seed(pos, faces, (0.65, 0.6))
    for robot in (-1, rf):
        ref = om.cvp(w, vc, f, sp, robot_face=robot, invalid=invalid)
        got = api.CVPMeshPlanner(mm).waveFrontPropagation(f, sp, robot)
        assert got["outcome"] == ref["outcome"]
        check_cvp(got, ref)
    # unreachable robot -> NO_PATH_FOUND (cvp:912-918)
    vc2 = vc.copy(); vc2[(pos[:, 0] > 6.0) & (pos[:, 0] < 6.35)] = 3.0
    w2 = om.edge_weights(vc2, ed, 1.0)
    mm.setCosts(vc2, w2, invalid)
    ref = om.cvp(w2, vc2, f, sp, robot_face=rf, invalid=invalid)
    got = api.CVPMeshPlanner(mm).waveFrontPropagation(f, sp, rf)
    assert ref["outcome"] == 54 and got["outcome"] == 54
    check_cvp(got, ref)
    mm.close()


@pytest.mark.parametrize("n,factor", [(300, 1.0), (300, 2.0), (200, 3.0)])
def test_cvp_cost_weighted_non_causal(api, oracle_mod, n, factor):
    """edge_cost_factor > 0 makes the 'edge lengths' non-geometric: large non-causal back-steps,
    |t0a| > 1 (acos NaN in the reference), nested cascades below the water line."""
    rng = np.random.default_rng(5)
    costs = lambda pos: np.where(rng.random(pos.shape[0]) < 0.04, 1.2, rng.random(pos.shape[0]) * 0.7)
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, n, True, costs=costs, factor=factor)
    v, f, sp = centre_seed(pos, faces, (0.3, 0.35))
    for x in faces[f]:
        vc[x] = 0.1
    w = om.edge_weights(vc, ed, factor)
    mm.setCosts(vc, w)
    ref = om.cvp(w, vc, f, sp)
    assert ref["max_backstep"] > 0.2          # the case really is strongly non-causal
    for cluster in (-1, 4):
        mm.set_tuning(0.3, cluster, 0)
        got = api.CVPMeshPlanner(mm).waveFrontPropagation(f, sp)
        assert got["rounds"] < 2 * om.V       # converged (watchdog not hit)
        check_cvp(got, ref)
    mm.close()


def test_cvp_batch_matches_single(api, oracle_mod):
    """batched potentials (config 4 shape, small): every field equals the oracle's"""
    from mesh_navigation_b200 import synth
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, 100, True)
    goals = synth.batch_goal_vertices(om.V, 12, seed=1234)
    sfs = np.array([face_of_vertex(faces, g) for g in goals], np.uint32)
    sps = np.stack([pos[faces[f]].mean(0) for f in sfs]).astype(np.float32)
    mm.set_tuning(0.3, 1, 0)
    got = api.CVPMeshPlanner(mm).waveFrontPropagationBatch(sfs, sps)
    for i in range(len(goals)):
        ref = om.cvp(w, vc, int(sfs[i]), sps[i])
        assert rel_err(got["dist"][i], ref["dist"]).max() <= CVP_RTOL
    mm.close()


def test_large_mesh_properties(api, oracle_mod):
    """1M-vertex terrain (config 2): parity vs the oracle + size-independent properties."""
    pos, faces, om, mm, ed, vc, w = setup(api, oracle_mod, 1000, True)
    v, f, sp = centre_seed(pos, faces)
    got = api.CVPMeshPlanner(mm).waveFrontPropagation(f, sp)
    ref = om.cvp(w, vc, f, sp)
    assert rel_err(got["dist"], ref["dist"]).max() <= CVP_RTOL
    d = got["dist"]
    assert np.isfinite(d).all() and got["settled"] >= om.V - 3
    eu = np.linalg.norm(pos - sp, axis=1)
    assert (d >= eu - 1e-4).all()                                    # a geodesic is never shorter than the chord
    p = got["pred"]; nz = p != np.arange(om.V)
    assert nz.sum() == om.V - 3                                      # everything but the 3 seeds has a predecessor
    gd = api.DijkstraMeshPlanner(mm).dijkstra(v)
    rd = om.dijkstra(w, vc, v)
    assert (gd["dist"].view(np.uint32) == rd["dist"].view(np.uint32)).all() and (gd["pred"] == rd["pred"]).all()
    assert (gd["dist"][gd["pred"][nz]] <= gd["dist"][nz]).all()
    # triangle inequality along every edge: |d(a) - d(b)| <= w(a,b)
    a, b = om.edges[:, 0], om.edges[:, 1]
    assert (np.abs(gd["dist"][a] - gd["dist"][b]) <= w + 2e-5).all()      # one float ulp at d ~ 100 m is 7.6e-6
    mm.close()
