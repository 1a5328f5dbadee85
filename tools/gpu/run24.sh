#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_run24_pytest.log
tail -3 gpurun_out/r02_run24_pytest.log
python - <<'PY'
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams
for kw in (dict(theta_m=0.2), dict(theta_m=0.2, bndry_lyr=0.5), dict(theta_m=0.2, bndry_lyr=0.8)):
    h = lib.Handle(ArtParams.default(**kw), 0)
    h.run_device(1769, 1, 300000)
    st = h.run_device(1769, 1, 1000000)["stats"]
    print(kw, "ms_total %.1f ms_sampler %.1f rays %d n_rhs/ray %.1f scanned/ray %.2f cond_eval/ray %.1f sampler cond/attempt %.1f" % (st["ms_total"], st["ms_sampler"], st["n_final_photon"], st["n_rhs"] / st["n_rays"], st["n_scanned_steps"] / st["n_rays"], st["n_cond_eval"] / st["n_rays"], st["n_sampler_cond"] / st["n_attempts"]))
    h.close()
PY
