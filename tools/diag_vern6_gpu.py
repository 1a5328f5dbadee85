"""GPU vs oracle at the reference tolerances, per integrator: how many rays take the identical accept/reject sequence, and the
distribution of the differences in the final state (python tools/diag_vern6_gpu.py [n])."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adiabatic_raytracer_b200 import lib  # noqa: E402
from adiabatic_raytracer_b200.params import ArtParams  # noqa: E402
from oracle import loader as oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for ig in (0, 1):
    for kw in (dict(theta_m=0.2), dict(theta_m=0.2, flat=1, omega_pul=2 * math.pi)):
        p = ArtParams.default(integrator=ig, **kw)
        h = lib.Handle(p, 0)
        s = oracle.sample(p, 5, 1, n)
        w = oracle.weights(p, s["x"], s["v_loc"], s["v_inf"])
        for sp, mc, bs in ((1, -1, 1), (0, 100000, -1)):
            k0 = w["k_init"] * bs
            args = (np.full(n, sp, np.int32), s["x"], k0, w["erg_inf"], np.zeros(n), -np.ones(n), bs, mc)
            a = h.propagate(*args)
            b = oracle.propagate(p, *args)
            same = (a["status"] == b["status"]) & (a["n_cross"] == b["n_cross"])
            steps = same & (a["counters"][:, 0] == b["counters"][:, 0]) & (a["counters"][:, 1] == b["counters"][:, 1])
            dx = np.linalg.norm(a["x_f"] - b["x_f"], axis=0) / np.linalg.norm(b["x_f"], axis=0)
            dk = np.linalg.norm(a["k_f"] - b["k_f"], axis=0) / np.linalg.norm(b["k_f"], axis=0)
            q = lambda v: " ".join(f"{np.percentile(v, pc):.1e}" for pc in (50, 75, 90, 99, 100))
            print(f"integrator {ig} {kw} species {sp}: same outcome {same.mean():.4f}, identical step counts {steps.mean():.4f}; "
                  f"dx p50/75/90/99/max {q(dx[same])}; dk {q(dk[same])}; within 1e-6: {(dk[same] < 1e-6).mean():.3f}; "
                  f"acc/rej/rhs gpu {a['counters'][:, :3].mean(axis=0).round(1)} oracle {b['counters'][:, :3].mean(axis=0).round(1)}",
                  flush=True)
        h.close()
