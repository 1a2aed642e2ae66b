import sys, numpy as np
sys.path.insert(0,'.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner
n=int(sys.argv[1]); ng=int(sys.argv[2])
pos,faces=synth.grid_mesh(n,n,terrain=True)
mm=MeshMap(pos,faces); ed=mm.edgeDistances(); mm.setCosts(np.zeros(mm.V,np.float32), ed)
goals=synth.batch_goal_vertices(mm.V, ng, seed=1234)
gi,gj=np.minimum(goals%n,n-2), np.minimum(goals//n,n-2)
sfs=(2*(gj*(n-1)+gi)).astype(np.uint32); sps=pos[faces[sfs]].mean(1).astype(np.float32)
mm.set_tuning(0.0,1,0)
for rep in range(2):
    B=CVPMeshPlanner(mm).waveFrontPropagationBatch(sfs, sps); print("batch ms", B['kernel_ms'], "recomp/V", B['recomputes']/ng/mm.V, "rounds", B['rounds']/ng)
