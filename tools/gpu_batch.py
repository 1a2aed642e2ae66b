import sys, time, numpy as np, torch
sys.path.insert(0,'.')
from mesh_navigation_b200 import synth
from mesh_navigation_b200.api import MeshMap, CVPMeshPlanner
n=int(sys.argv[1]); 
pos,faces=synth.grid_mesh(n,n,terrain=True)
mm=MeshMap(pos,faces); ed=mm.edgeDistances(); mm.setCosts(np.zeros(mm.V,np.float32), ed)
goals=synth.batch_goal_vertices(mm.V, 1024, seed=1234)
gi,gj=goals%n, goals//n; gi=np.minimum(gi,n-2); gj=np.minimum(gj,n-2)
sfs=(2*(gj*(n-1)+gi)).astype(np.uint32); sps=pos[faces[sfs]].mean(1).astype(np.float32)
for spec in sys.argv[2:]:
    cs,ng,delta=spec.split(','); cs=int(cs); ng=int(ng); delta=float(delta)
    mm.set_tuning(delta, cs, 0)
    out=torch.empty((ng,mm.V),dtype=torch.float32,device='cuda')
    mm.use_device_pointers(True)
    for rep in range(2):
        t=time.time(); mm.cvp_batch_dev(sfs[:ng], sps[:ng], 1.0, out.data_ptr()); torch.cuda.synchronize(); dt=time.time()-t
    mm.use_device_pointers(False)
    st=mm.stats()
    print(f"n={n} cluster={cs} goals={ng} delta={delta}: {dt*1e3:.1f} ms -> {ng/dt:.1f} plans/s, {ng*mm.V/dt/1e6:.1f} Mrelax/s, kernel {st['kernel_ms']:.1f} ms rounds/plan {st['rounds']/ng:.0f} recomp/V {st['recomputes']/ng/mm.V:.2f}", flush=True)
