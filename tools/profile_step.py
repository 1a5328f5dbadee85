"""One small pipeline pass for profiling (ncu): python tools/profile_step.py [events]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
h = lib.Handle(ArtParams.default(theta_m=0.2), 0)
out = h.run_device(1769, 1, n)
st = out["stats"]
print({k: st[k] for k in ("n_rays", "n_acc", "n_rej", "n_rhs", "n_cond", "ms_sampler", "ms_tracer", "ms_tree", "ms_total", "n_launches")})
h.close()
