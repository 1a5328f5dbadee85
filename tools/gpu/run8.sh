#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_run8_pytest.log
tail -4 gpurun_out/r02_run8_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run8_bench.json 2> gpurun_out/r02_run8_bench.err; head -c 200 gpurun_out/r02_run8_bench.json
python - <<'PY'
import time, numpy as np
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams
for kw in (dict(theta_m=0.2), dict(theta_m=0.2, bndry_lyr=0.5)):
    h = lib.Handle(ArtParams.default(**kw), 0)
    h.run_device(1769, 1, 300000)
    st = h.run_device(1769, 1, 1000000)["stats"]
    print(kw, "ms_total %.1f rays %d n_rhs/ray %.1f" % (st["ms_total"], st["n_final_photon"], st["n_rhs"] / st["n_rays"]))
    h.close()
PY
