import numpy as np, ctypes as C, sys, time
sys.path.insert(0,'.')
from oracle import oracle as O
from tests.util import *
L=C.CDLL('tools/libsimband.so'); vp=C.c_void_p
L.sim_cvp_band.argtypes=[C.c_uint32,C.c_uint32,vp,vp,C.c_uint32,vp,vp,vp,vp,C.c_uint32,vp,C.c_double,C.c_double,C.c_int,vp,vp]
def p(a): return None if a is None else a.ctypes.data_as(vp)
def run(n, terrain, with_costs, factor, with_invalid, delta=0.3, seedfrac=(0.3,0.35)):
    rng = np.random.default_rng(5)
    pos, faces = mesh_case(n, terrain)
    m = O.OracleMesh(pos, faces); ed = m.edge_distances()
    vc = np.where(rng.random(m.V) < 0.04, 1.2, rng.random(m.V) * 0.7).astype(np.float32) if with_costs else np.zeros(m.V, np.float32)
    invalid = (rng.random(m.V) < 0.005).astype(np.uint8) if with_invalid else None
    v, f, sp = centre_seed(pos, faces, seedfrac)
    for x in faces[f]:
        vc[x] = 0.1 if with_costs else 0
        if invalid is not None: invalid[x] = 0
    w = m.edge_weights(vc, ed, factor)
    ref = m.cvp(w, vc, f, sp, invalid=invalid)
    out = np.empty(m.V, np.float32); st = np.zeros(4)
    t=time.time()
    L.sim_cvp_band(m.V, m.F, p(m.faces), p(m.edges), m.E, p(m.pos), p(w), p(vc), p(invalid), f, p(sp), 1.0, delta, 0, p(out), p(st))
    fr, fg = np.isfinite(ref['dist']), np.isfinite(out)
    both = fr & fg
    print(f"n={n} terr={terrain} costs={with_costs} f={factor} inv={with_invalid}: rounds={int(st[0])} watchdog={int(st[2])} recomp/V={st[1]/m.V:.2f} reached ref/sim {fr.sum()}/{fg.sum()} neq={(out[both]!=ref['dist'][both]).sum()} backsteps={ref['backsteps']} maxback={ref['max_backstep']:.3f} t={time.time()-t:.1f}s", flush=True)
if __name__ == "__main__":
    run(100, False, False, 0.0, False)
    run(140, True, True, 1.0, True)
    run(140, True, True, 1.0, False)
    run(140, True, True, 0.0, False)
    run(300, True, True, 2.0, True)
    run(1000, True, False, 0.0, False)

def detail(n, terrain, factor, delta=0.3, seedfrac=(0.3,0.35)):
    rng = np.random.default_rng(5)
    pos, faces = mesh_case(n, terrain)
    m = O.OracleMesh(pos, faces); ed = m.edge_distances()
    vc = np.where(rng.random(m.V) < 0.04, 1.2, rng.random(m.V) * 0.7).astype(np.float32)
    v, f, sp = centre_seed(pos, faces, seedfrac)
    for x in faces[f]: vc[x] = 0.1
    w = m.edge_weights(vc, ed, factor)
    ref = m.cvp(w, vc, f, sp)
    out = np.empty(m.V, np.float32); st = np.zeros(4)
    L.sim_cvp_band(m.V, m.F, p(m.faces), p(m.edges), m.E, p(m.pos), p(w), p(vc), None, f, p(sp), 1.0, delta, 0, p(out), p(st))
    fin = np.isfinite(ref['dist'])
    rel = np.abs(out[fin]-ref['dist'][fin])/np.maximum(ref['dist'][fin],1e-30)
    print("factor",factor,"maxrel",rel.max(),"n>1e-4",(rel>1e-4).sum(),"of",fin.sum(), "mean rel", rel.mean())
    idx = np.where(fin)[0][np.argsort(ref['dist'][fin])]
    bad = [c for c in idx if out[c]!=ref['dist'][c]][:3]
    for c in bad:
        print(" first diff v",c,"ref",ref['dist'][c],"sim",out[c],"pred",ref['pred'][c],"cut",ref['cutting_face'][c])
        for fc in np.where((faces==c).any(1))[0]:
            print("    face",fc,[(int(x),float(ref['dist'][x]),float(out[x])) for x in faces[fc]])
if __name__ == "__main__" and len(sys.argv)>1 and sys.argv[1] != "x":
    detail(140, True, float(sys.argv[1]))
if __name__ == "__main__" and len(sys.argv)>1 and sys.argv[1] == "x":
    for (n,fac) in ((300,2.0),(300,1.0),(400,3.0)):
        rng = np.random.default_rng(5)
        pos, faces = mesh_case(n, True)
        m = O.OracleMesh(pos, faces); ed = m.edge_distances()
        vc = np.where(rng.random(m.V) < 0.04, 1.2, rng.random(m.V) * 0.7).astype(np.float32)
        v, f, sp = centre_seed(pos, faces, (0.3,0.35))
        for x in faces[f]: vc[x] = 0.1
        w = m.edge_weights(vc, ed, fac)
        ref = m.cvp(w, vc, f, sp)
        out = np.empty(m.V, np.float32); st = np.zeros(4)
        for lv in (2,3,4,6,8,12):
          L.sim_cvp_band(m.V, m.F, p(m.faces), p(m.edges), m.E, p(m.pos), p(w), p(vc), None, f, p(sp), 1.0, 0.3, lv, p(out), p(st))
          fin = np.isfinite(ref['dist'])
          rel = np.abs(out[fin]-ref['dist'][fin])/np.maximum(ref['dist'][fin],1e-30)
          print("levels",lv,"n",n,"factor",fac,"neq",(out[fin]!=ref['dist'][fin]).sum(),"maxrel",rel.max(),"n>1e-4",(rel>1e-4).sum(),"of",fin.sum())
