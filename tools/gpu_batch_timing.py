"""Dev script (GPU box): phase cycles of the batch round loop (library built with -DMNB_BATCH_TIMING, build/variants/timing.so)"""
import os, sys
os.environ["MNB_PHASE_TIMING"] = "1"
sys.path.insert(0, '.')
from mesh_navigation_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), "build/variants/timing.so")
import numpy as np
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner
n = int(sys.argv[1]); ng = int(sys.argv[2]); cs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pos, faces = synth.grid_mesh(n, n, terrain=True, seed=42)
mm = MeshMap(pos, faces); mm.setCosts(np.zeros(mm.V, np.float32), mm.edgeDistances())
goals = synth.batch_goal_vertices(mm.V, ng, seed=1234)
gi, gj = np.minimum(goals % n, n - 2), np.minimum(goals // n, n - 2)
sfs = (2 * (gj * (n - 1) + gi)).astype(np.uint32); sps = pos[faces[sfs]].mean(1).astype(np.float32)
mm.set_tuning(0.3, cs, 0)
B = CVPMeshPlanner(mm).waveFrontPropagationBatch(sfs, sps)
print("batch ms", B['kernel_ms'], "recomp/V", B['recomputes'] / ng / mm.V, "rounds", B['rounds'] / ng, file=sys.stderr)
