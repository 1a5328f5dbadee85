#!/bin/bash
mkdir -p gpurun_out
timeout 100 python - <<'PY'
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams
h = lib.Handle(ArtParams.default(theta_m=0.2, bndry_lyr=0.5), 0)
st = h.run_device(1769, 1, 100000)["stats"]
print("100000 events: ms_total %.1f status %s" % (st["ms_total"], st["n_status"]))
for n in (1000000, 5000000):
    st = h.run_device(1769, 1, n)["stats"]
    print(n, "events: ms_total %.1f = %.1f ms per 1e6, launches %d, status %s" % (st["ms_total"], st["ms_total"] * 1e6 / n, st["n_launches"], st["n_status"]))
PY
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_run26_pytest.log
tail -3 gpurun_out/r02_run26_pytest.log
