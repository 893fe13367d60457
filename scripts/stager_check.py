"""dev: pageable 120 MB particle upload through mms_push_particles, with and without the HostStager (wall clock incl. a stream sync)."""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import megamol_b200 as mm
    from megamol_b200 import synth
    xyz, box = synth.lj_fluid(10_000_000)
    s = mm.Surf(0)
    s.set_grid((0, 0, 0), (box,) * 3, (512, 512, 512), (True,) * 3)
    s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
    ts = []
    for it in range(6):
        s.clear_particles()
        t0 = time.perf_counter()
        s.push_particles([dict(vtx=xyz, vtx_type=1, count=len(xyz), global_radius=0.5)])
        s.L.mms_synchronize(s.h)
        ts.append((time.perf_counter() - t0) * 1e3)
        s.compute_density()
    mn, mx = s.density_range()
    print(sys.argv[1], "push+sync ms:", [round(t, 2) for t in ts], "range", mn, mx)
    s.close()
else:
    for tag, env in (("plain", {"MMS_NO_STAGER": "1"}), ("staged", {})):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, tag], env=e)
