import sys, time, numpy as np
sys.path.insert(0,'.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, InflationLayer
from tests.util import disc_lethals
for n in [int(x) for x in sys.argv[1:]]:
    pos,faces=synth.grid_mesh(n,n,terrain=True)
    mm=MeshMap(pos,faces)
    L=mm.computeLayers(); print(n,"layers ms",L['kernel_ms'],"lethal",int((L['lethal_mask']!=0).sum()), flush=True)
    le=np.union1d(np.where(L['lethal_mask']!=0)[0], disc_lethals(pos, 1000, 0.3)).astype(np.uint32)
    t=time.time(); I=InflationLayer(mm).waveCostInflation(le); print(n,"inflate ms",I['kernel_ms'],"wall",time.time()-t,"rounds",I['rounds'],"labelled",int(np.isfinite(I['dist']).sum()),"recomputes",I['recomputes'], flush=True)
    mm.close()
