import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
from oracle import bindings as ob
xf = load_package()
DT = np.float32(1/3000)
def settings(mod, damped):
    kw = dict(energy=7, poisson=0.5)
    if damped: kw.update(damping=0.005, rayleigh=3, pbd_damping=0.03)
    st = mod.make_settings(**kw)
    if damped:
        sdt=1/3000; k=(20/100)/31
        st.volumeAndTimeCorrectedPbdDamping=(1-(1-0.03)**(1000*sdt))*6*k*k
        st.amortizedVolumeAndTimeCorrectedPbdDamping=(1-(1-0.03)**(8000*sdt))*6*k*k
        st.timeCorrectedDrag=1-(1-0.002)**(1000*sdt)
    return st
for cells in (16, 55):
    for damped in (False, True):
        nodes, idx, hint = xf.GenerateTetBlock(cells, cells)
        geo = xf.GeoLinear3dCuda(nodes, idx, color_hint=hint)
        st = settings(xf, damped)
        s0 = geo.stats(st)
        for f in range(3):
            geo.Substep(st, DT, 100); st.tickId += 100
            s = geo.stats(st)
            print(cells, damped, "K %.3e released %.3e dev %.3e kernel %s" % (s["kinetic"], s0["gravitational"]-s["gravitational"], s["deviatoric"]-s0["deviatoric"], geo.info()["lastKernel"]), flush=True)
        if cells == 55:
            X, V, w = geo.get_state()
            orc = ob.OracleScene(nodes, idx); orc.set_order(geo.get_order()); orc.set_state(X, V, w)
            ost = settings(ob, damped); ost.tickId = st.tickId
            e = orc.energy(ost)
            print("  oracle energy of the same state: K %.3e grav %.3e dev %.3e" % (e[0], e[1], e[2]))
            geo.Substep(st, DT, 2); orc.substep(ost, DT, 2)
            Xg, Vg, wg = geo.get_state(); Xo, Vo, wo = orc.get_state()
            print("  teacher-forced 2 substeps at 1M, damped=%s: X equal %s V equal %s w equal %s maxdX %.3e" % (damped, np.array_equal(Xg,Xo), np.array_equal(Vg,Vo), np.array_equal(wg,wo), np.abs(Xg-Xo).max()), flush=True)
        geo.close()
