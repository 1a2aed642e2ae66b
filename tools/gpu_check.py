"""Dev script (GPU box): parity + timing of the CUDA wavefronts vs the oracle."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle as O
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, DijkstraMeshPlanner, CVPMeshPlanner

def case(n, terrain, cluster, delta, threads=512):
    pos, faces = synth.grid_mesh(n, n, terrain=terrain)
    om = O.OracleMesh(pos, faces)
    mm = MeshMap(pos, faces)
    mm.set_tuning(delta, cluster, threads)
    assert (mm.edges() == om.edges).all()
    ed = om.edge_distances()
    ged = mm.edgeDistances()
    print("edge dist bit-equal:", bool((ged.view(np.uint32) == ed.view(np.uint32)).all()))
    vc = np.zeros(om.V, np.float32)
    mm.setCosts(vc, ed)
    seed = synth.nearest_vertex(pos, [n * 0.05, n * 0.05, float(pos[:, 2].mean())])
    sf = int(np.where((faces == seed).any(1))[0][0]); sp = pos[faces[sf]].mean(0).astype(np.float32)
    # dijkstra
    ref = om.dijkstra(ed, vc, seed)
    t = time.time(); got = DijkstraMeshPlanner(mm).dijkstra(seed); wall = time.time() - t
    print(f"[dijkstra n={n} cs={cluster} d={delta}] dist!= {(got['dist'].view(np.uint32) != ref['dist'].view(np.uint32)).sum()} "
          f"pred!= {(got['pred'] != ref['pred']).sum()} rounds={got['rounds']} recomp/V={got['recomputes']/om.V:.2f} "
          f"kernel_ms={got['kernel_ms']:.3f} wall={wall*1e3:.1f}ms cpu={ref['seconds']*1e3:.1f}ms")
    ref = om.cvp(ed, vc, sf, sp)
    t = time.time(); got = CVPMeshPlanner(mm).waveFrontPropagation(sf, sp); wall = time.time() - t
    rel = np.abs(got['dist'] - ref['dist']) / np.maximum(ref['dist'], 1e-30)
    rel[~np.isfinite(ref['dist'])] = 0
    print(f"[cvp n={n} cs={cluster} d={delta}] dist!= {(got['dist'].view(np.uint32) != ref['dist'].view(np.uint32)).sum()} "
          f"maxrel={rel.max():.3e} pred!= {(got['pred'] != ref['pred']).sum()} cut!= {(got['cutting_face'] != ref['cutting_face']).sum()} "
          f"dir maxabs={np.abs(got['direction']-ref['direction']).max():.2e} rounds={got['rounds']} recomp/V={got['recomputes']/om.V:.2f} "
          f"settled={got['settled']} kernel_ms={got['kernel_ms']:.3f} wall={wall*1e3:.1f}ms cpu={ref['seconds']*1e3:.1f}ms", flush=True)
    mm.close()

if __name__ == "__main__":
    for spec in sys.argv[1:]:
        n, terrain, cluster, delta = spec.split(',')
        case(int(n), bool(int(terrain)), int(cluster), float(delta))
