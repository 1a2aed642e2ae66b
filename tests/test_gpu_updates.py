"""GPU parity tests of the incremental-update entry points: MeshMap::layerChanged +
updateEdgeWeights, MaxCombinationLayer::onInputChanged, InflationLayer::onInputChanged -- through the C ABI, against the
oracle's restatements.  (Also replayed on the CPU interpreter of the kernels by tests/test_emu_kernels.py.)"""
import numpy as np
import pytest

from tests.util import centre_seed, disc_lethals, mesh_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mesh_navigation_b200 import api as A
    return A


def _plans(api, mm, pos, faces):
    v, f, sp = centre_seed(pos, faces, (0.4, 0.6))
    c = api.CVPMeshPlanner(mm, cost_limit=2.0).waveFrontPropagation(f, sp)
    d = api.DijkstraMeshPlanner(mm, cost_limit=2.0).dijkstra(v)
    return c, d


@pytest.mark.parametrize("n,terrain,factor,frac", [(60, True, 1.5, 0.02), (90, False, 3.0, 0.2), (48, True, 0.0, 0.1)])
def test_layer_changed_matches_full_reinstall(api, oracle_mod, n, terrain, factor, frac):
    """after mnb_update_vertex_costs the device holds exactly what the oracle's layerChanged + updateEdgeWeights produce,
    and the planners (whose corner / ELL / adjacency tables were patched in place) give bit-identical fields to a map
    freshly installed with those arrays"""
    O = oracle_mod
    pos, faces = mesh_case(n, terrain)
    om = O.OracleMesh(pos, faces)
    rng = np.random.default_rng(11)
    ed = om.edge_distances()
    vc = (rng.random(om.V) * 0.8).astype(np.float32)
    w = om.edge_weights(vc, ed, factor)
    mm = api.MeshMap(pos, faces)
    mm.setCosts(vc, w)
    changed = rng.choice(om.V, max(1, int(frac * om.V)), replace=False).astype(np.uint32)
    new = (rng.random(changed.size) * 1.2).astype(np.float32)
    new[::17] = np.inf                                      # +inf endpoint -> +inf weight (mesh_map.cpp:598)
    layer = np.full(om.V, np.nan, np.float32); layer[changed] = new
    # oracle
    vc_ref = vc.copy(); O.layer_changed(layer, 0.0, changed, vc_ref)
    w_ref = w.copy(); om.update_edge_weights(vc_ref, ed, factor, changed, w_ref)
    # device, per-changed-vertex values; duplicates in the list are allowed
    ch_dup = np.concatenate([changed, changed[:5]]); new_dup = np.concatenate([new, new[:5]])
    mm.layerChanged(ch_dup, new_dup, factor)
    gvc, gw = mm.costs()
    assert (gvc.view(np.uint32) == vc_ref.view(np.uint32)).all()
    assert (gw.view(np.uint32) == w_ref.view(np.uint32)).all()
    if factor == 0.0:
        assert (gw.view(np.uint32) == w.view(np.uint32)).all(), "zero factor must leave the weights alone (mesh_map.cpp:568)"
    got_c, got_d = _plans(api, mm, pos, faces)
    fresh = api.MeshMap(pos, faces)
    fresh.setCosts(vc_ref, w_ref)
    ref_c, ref_d = _plans(api, fresh, pos, faces)
    for k in ("dist", "direction"):
        assert (got_c[k].view(np.uint32) == ref_c[k].view(np.uint32)).all(), k
    assert (got_c["pred"] == ref_c["pred"]).all() and (got_c["cutting_face"] == ref_c["cutting_face"]).all()
    assert (got_d["dist"].view(np.uint32) == ref_d["dist"].view(np.uint32)).all() and (got_d["pred"] == ref_d["pred"]).all()
    # and against the oracle itself
    v, f, sp = centre_seed(pos, faces, (0.4, 0.6))
    oc = om.cvp(w_ref, vc_ref, f, sp, cost_limit=2.0)
    assert (got_c["dist"].view(np.uint32) == oc["dist"].view(np.uint32)).all()
    # the V-sized-map form gives the same result
    mm2 = api.MeshMap(pos, faces); mm2.setCosts(vc, w)
    mm2.layerChanged(changed, layer, factor, by_vertex=True, default_value=0.0)
    g2, w2 = mm2.costs()
    assert (g2.view(np.uint32) == vc_ref.view(np.uint32)).all() and (w2.view(np.uint32) == w_ref.view(np.uint32)).all()
    for m in (mm, mm2, fresh):
        m.close()


def test_max_combination_update(api, oracle_mod):
    O = oracle_mod
    pos, faces = mesh_case(50, True)
    V = pos.shape[0]
    rng = np.random.default_rng(3)
    la = rng.random(V).astype(np.float32)
    lb = np.full(V, np.nan, np.float32); idx = rng.choice(V, V // 3, replace=False); lb[idx] = (rng.random(idx.size) * 2).astype(np.float32)
    lc = np.full(V, np.nan, np.float32)
    leth_a = (la > 0.9).astype(np.uint8); leth_b = (np.nan_to_num(lb) > 1.0).astype(np.uint8)
    changed = rng.choice(V, V // 4, replace=False).astype(np.uint32)
    ref_c = np.full(V, -1.0, np.float32); ref_l = np.full(V, 7, np.uint8)
    O.max_combination_update([la, lb, lc], [0.0, 0.0, 0.25], [leth_a, leth_b, None], changed, ref_c, ref_l)
    mm = api.MeshMap(pos, faces)
    got_c = np.full(V, -1.0, np.float32); got_l = np.full(V, 7, np.uint8)
    mm.maxCombinationUpdate([la, lb, lc], [0.0, 0.0, 0.25], [leth_a, leth_b, None], changed, got_c, got_l)
    assert (got_c.view(np.uint32) == ref_c.view(np.uint32)).all() and (got_l == ref_l).all()
    untouched = np.setdiff1d(np.arange(V), changed)
    assert (got_c[untouched] == -1.0).all() and (got_l[untouched] == 7).all()
    mm.close()


def test_inflation_on_input_changed_chain(api, oracle_mod):
    """dynamic obstacle cycle: obstacles appear, move, disappear.  Every cycle: re-inflation + update set,
    final = max(static, riskiness) on the update set, layerChanged on the planner map; the planner then sees exactly the
    field a from-scratch install would give."""
    O = oracle_mod
    n = 72
    pos, faces = mesh_case(n, True)
    om = O.OracleMesh(pos, faces)
    V = om.V
    ed = om.edge_distances()
    rng = np.random.default_rng(5)
    static = (rng.random(V) * 0.3).astype(np.float32)
    factor = 2.0
    mm = api.MeshMap(pos, faces)
    vc_ref = static.copy(); w_ref = om.edge_weights(vc_ref, ed, factor)
    mm.setCosts(vc_ref, w_ref)
    infl = api.InflationLayer(mm)
    final_ref = static.copy(); final_got = static.copy()
    old_cost = None
    for cyc, (discs, seed) in enumerate([(5, 1), (5, 2), (0, 0), (9, 4)]):
        lethals = disc_lethals(pos, discs, 0.25, seed=seed) if discs else np.empty(0, np.uint32)
        ref = om.inflation(ed, lethals)
        upd_ref = O.inflation_update_set(ref["cost"], old_cost)
        got = infl.onInputChanged(lethals)
        assert (got["changed"] == upd_ref).all(), f"cycle {cyc}: update set differs"
        assert np.array_equal(np.isnan(got["cost"]), np.isnan(ref["cost"]))
        fin = ~np.isnan(ref["cost"])
        assert np.allclose(got["cost"][fin], ref["cost"][fin], rtol=1e-5, atol=0)
        old_cost = ref["cost"]
        # final combination layer on the update set, from the DEVICE's riskiness so that both sides see identical inputs
        O.max_combination_update([static, got["cost"]], [0.0, 0.0], [None, None], upd_ref, final_ref, None)
        mm.maxCombinationUpdate([static, got["cost"]], [0.0, 0.0], None, got["changed"], final_got, None)
        assert (final_got.view(np.uint32) == final_ref.view(np.uint32)).all()
        O.layer_changed(final_ref, 0.0, upd_ref, vc_ref)
        om.update_edge_weights(vc_ref, ed, factor, upd_ref, w_ref)
        mm.layerChanged(got["changed"], final_got, factor, by_vertex=True)
        gvc, gw = mm.costs()
        assert (gvc.view(np.uint32) == vc_ref.view(np.uint32)).all() and (gw.view(np.uint32) == w_ref.view(np.uint32)).all()
        v, f, sp = centre_seed(pos, faces, (0.5, 0.5))
        oc = om.cvp(w_ref, vc_ref, f, sp, cost_limit=0.9)
        gc = api.CVPMeshPlanner(mm, cost_limit=0.9).waveFrontPropagation(f, sp)
        assert (gc["dist"].view(np.uint32) == oc["dist"].view(np.uint32)).all(), f"cycle {cyc}: plan differs"
    mm.close()
