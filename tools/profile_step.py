"""Pipeline passes for profiling / tuning: python tools/profile_step.py [events] [repeats]
Prints the stage times of every pass (the first pass of a process includes device allocations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = lib.Handle(ArtParams.default(theta_m=0.2), 0)
keys = ("n_rays", "n_acc", "n_rej", "n_rhs", "n_cond", "ms_sampler", "ms_tracer", "ms_total", "n_launches")
for r in range(reps):
    st = h.run_device(1769, 1, n)["stats"]
    if reps == 1:
        print({k: st[k] for k in keys})
    else:
        print(f"pass {r}: sampler {st['ms_sampler']:.1f} tracer {st['ms_tracer']:.1f} total {st['ms_total']:.1f} ms, launches {st['n_launches']}")
h.close()
