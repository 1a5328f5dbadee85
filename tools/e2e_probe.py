"""Where does the end-to-end time of art_run go?  python tools/e2e_probe.py [events]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
h = lib.Handle(ArtParams.default(theta_m=0.2), 0)
pin = lib.PinnedRows(n * 3 + 1024, 13)
host = pin.array
for r in range(4):
    t0 = time.perf_counter()
    out = h.run_device(1769, 1, n)
    t1 = time.perf_counter()
    rows = h.fetch_rows(out["n_rows"])
    t2 = time.perf_counter()
    o2 = h.run(1769, 1, n, rows_out=host)
    t3 = time.perf_counter()
    print(f"pass {r}: run_device {1e3*(t1-t0):.1f} ms (device {out['stats']['ms_total']:.1f}), fetch_rows {1e3*(t2-t1):.1f} ms, "
          f"run(rows_out) {1e3*(t3-t2):.1f} ms, rows {out['n_rows']}")
h.close()
