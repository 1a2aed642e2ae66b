import sys, numpy as np
sys.path.insert(0,'.')
from oracle import oracle as O
from mesh_navigation_b200 import api as A
from tests.util import *
rng = np.random.default_rng(5)
costs = lambda pos: np.where(rng.random(pos.shape[0]) < 0.04, 1.2, rng.random(pos.shape[0]) * 0.7)
pos, faces = mesh_case(140, True)
invalid = (rng.random(pos.shape[0]) < 0.005).astype(np.uint8)
om = O.OracleMesh(pos, faces); mm = A.MeshMap(pos, faces)
ed = om.edge_distances(); vc = costs(pos).astype(np.float32)
v, f, sp = centre_seed(pos, faces, (0.3, 0.35))
for x in faces[f]:
    invalid[x] = 0; vc[x] = 0.1
w = om.edge_weights(vc, ed, 1.0); mm.setCosts(vc, w, invalid)
rv, rf, _ = centre_seed(pos, faces, (0.65, 0.6))
for cl in (-1, 8):
  mm.set_tuning(0.3, cl, 0)
  for robot in (-1, rf):
    ref = om.cvp(w, vc, f, sp, robot_face=robot, invalid=invalid)
    got = A.CVPMeshPlanner(mm).waveFrontPropagation(f, sp, robot)
    fr, fg = np.isfinite(ref['dist']), np.isfinite(got['dist'])
    print("cluster",cl,"robot",robot,"outcome",got['outcome'],ref['outcome'],"reached ref",fr.sum(),"got",fg.sum(),"only ref",(fr&~fg).sum(),"only got",(fg&~fr).sum(), "neq among both", (got['dist'][fr&fg]!=ref['dist'][fr&fg]).sum(), "rounds", got['rounds'])
    bad = np.where(fr!=fg)[0]
    for c in bad[:4]:
        print("  v",c,"ref",ref['dist'][c],"got",got['dist'][c],"cost",vc[c],"inv",invalid[c], "refpred", ref['pred'][c], "d(pred)", ref['dist'][ref['pred'][c]], "cut", ref['cutting_face'][c], faces[ref['cutting_face'][c]] if ref['cutting_face'][c]>=0 else None)
        fc=ref['cutting_face'][c]
        if fc>=0:
            for x in faces[fc]: print("     ",x,"ref",ref['dist'][x],"got",got['dist'][x],"cost",vc[x],"inv",invalid[x])
