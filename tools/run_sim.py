import numpy as np, ctypes as C, sys, time
sys.path.insert(0,'.')
from oracle import oracle as O
from mesh_navigation_b200 import synth
L=C.CDLL('tools/libsimband.so')
vp=C.c_void_p
L.sim_cvp_band.argtypes=[C.c_uint32,C.c_uint32,vp,vp,C.c_uint32,vp,vp,vp,vp,C.c_uint32,vp,C.c_double,C.c_double,C.c_int,vp,vp]
def p(a): return None if a is None else a.ctypes.data_as(vp)
def run(n,terrain,deltas,flagsets):
    pos,faces=synth.grid_mesh(n,n,terrain=terrain)
    m=O.OracleMesh(pos,faces); ed=m.edge_distances(); vc=np.zeros(m.V,np.float32)
    seed=synth.nearest_vertex(pos,[n*0.05,n*0.05,pos[:,2].mean()])
    sf=int(np.where((faces==seed).any(1))[0][0]); sp=pos[faces[sf]].mean(0).astype(np.float32)
    ref=m.cvp(ed,vc,sf,sp)
    print(f"n={n} terrain={terrain} oracle {ref['seconds']:.3f}s backsteps={ref['backsteps']} maxback={ref['max_backstep']:.4f}")
    for fl in flagsets:
        for dl in deltas:
            out=np.empty(m.V,np.float32); st=np.zeros(4)
            t=time.time()
            L.sim_cvp_band(m.V,m.F,p(m.faces),p(m.edges),m.E,p(m.pos),p(ed),p(vc),None,sf,p(sp),1.0,dl,fl,p(out),p(st))
            rel=np.abs(out-ref['dist'])/np.maximum(ref['dist'],1e-30)
            nbad=(rel>1e-4).sum(); 
            print(f"  flags={fl} delta={dl}: rounds={int(st[0])} recomputes/V={st[1]/m.V:.2f} maxrel={rel.max():.3e} n>1e-4={nbad} n!=:{(out!=ref['dist']).sum()} mean_rel={rel.mean():.2e} t={time.time()-t:.1f}s")
if __name__=="__main__":
    n=int(sys.argv[1]); terrain=int(sys.argv[2])
    run(n,bool(terrain),[float(x) for x in sys.argv[3].split(',')],[int(x) for x in sys.argv[4].split(',')])
