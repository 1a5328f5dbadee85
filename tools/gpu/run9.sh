#!/bin/bash
# ncu captures of the shipped build: launch list (time + DRAM bytes per launch) of one 2e6-event pass through a 2^19-slot pool
# (single stream under ncu), --set full of the steady-state stage / scan / detect launches and of the sampler + side kernels
mkdir -p gpurun_out
export ART_SIDE_STREAM=0 ART_POOL_SLOTS=524288
timeout 300 python tools/profile_step.py 2000000 1 > gpurun_out/r02_run9_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_final_launches.csv python tools/profile_step.py 2000000 1 > gpurun_out/r02_run9_ncu_list.log 2>&1
tail -1 gpurun_out/r02_run9_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_wf_stages|k_wf_scan|k_wf_detect' -s 450 -c 6 \
    -f -o gpurun_out/r02_final_steady python tools/profile_step.py 2000000 1 > gpurun_out/r02_run9_ncu_steady.log 2>&1
tail -1 gpurun_out/r02_run9_ncu_steady.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sample_events|k_wf_root|k_wf_tree|k_wf_inject' -s 0 -c 4 \
    -f -o gpurun_out/r02_final_side python tools/profile_step.py 2000000 1 > gpurun_out/r02_run9_ncu_side.log 2>&1
tail -1 gpurun_out/r02_run9_ncu_side.log
gzip -f gpurun_out/r02_final_launches.csv
unset ART_SIDE_STREAM ART_POOL_SLOTS
python - <<'PY'
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams
h = lib.Handle(ArtParams.default(theta_m=0.2, bndry_lyr=0.5), 0)
h.run_device(1769, 1, 200000)
st = h.run_device(1769, 1, 1000000)["stats"]
print("bndry_lyr=0.5:", {k: st[k] for k in ("ms_total", "ms_sampler", "ms_tracer", "n_attempts", "n_sampler_cond", "n_rays", "n_rhs", "n_cond_eval", "n_scanned_steps", "n_status")})
PY
