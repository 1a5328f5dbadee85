"""Tuning sweep over the library's environment knobs: python tools/env_sweep.py EVENTS "K=V,K=V" "K=V" ...
Each setting gets a fresh handle (knobs are read in art_create), one warm-up pass and three timed passes;
prints the best pass and a checksum of the rows (tuning knobs must not change results)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams

n = int(sys.argv[1])
for setting in sys.argv[2:] or [""]:
    kv = dict(s.split("=") for s in setting.split(",") if s)
    os.environ.update(kv)
    h = lib.Handle(ArtParams.default(theta_m=0.2), 0)
    best = None
    for r in range(4):
        out = h.run_device(1769, 1, n)
        st = out["stats"]
        if r and (best is None or st["ms_total"] < best["ms_total"]):
            best = st
    rows = h.fetch_rows(out["n_rows"])
    order = np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))
    chk = float(np.sum(rows[order, 2] * (np.arange(1, len(rows) + 1) % 7919)))
    print(f"{setting or 'defaults'}: sampler {best['ms_sampler']:.1f} tracer {best['ms_tracer']:.1f} total {best['ms_total']:.1f} ms, "
          f"launches {best['n_launches']}, rows {out['n_rows']}, f_inx {out['f_inx']}, chk {chk:.6f}", flush=True)
    h.close()
    for k in kv:
        del os.environ[k]
