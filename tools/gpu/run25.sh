#!/bin/bash
mkdir -p gpurun_out
ART_WF_TRACE=1 python - > gpurun_out/r02_bl_trace.txt 2>&1 <<'PY'
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams
h = lib.Handle(ArtParams.default(theta_m=0.2, bndry_lyr=0.5), 0)
st = h.run_device(1769, 1, 1000000)["stats"]
print({k: st[k] for k in ("ms_total", "ms_sampler", "ms_tracer", "n_rays", "n_acc", "n_rej", "n_rhs", "n_status", "n_launches")})
PY
grep -c "\[wf\]" gpurun_out/r02_bl_trace.txt; grep "\[wf\]" gpurun_out/r02_bl_trace.txt | awk 'NR%40==1' | cut -c1-200 | head -40; tail -2 gpurun_out/r02_bl_trace.txt
