#!/bin/bash
# ncu captures of the shipped build: launch list (time + DRAM bytes per launch) of one 2e6-event pass through a 2^19-slot pool
# (single stream under ncu), --set full of the steady-state stage / scan / detect launches and of the sampler + side kernels
mkdir -p gpurun_out
export ART_SIDE_STREAM=0 ART_POOL_SLOTS=524288
timeout 300 python tools/profile_step.py 2000000 1 > gpurun_out/r02_run16_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_final2_launches.csv python tools/profile_step.py 2000000 1 > gpurun_out/r02_run16_ncu_list.log 2>&1
tail -1 gpurun_out/r02_run16_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_wf_stages|k_wf_scan|k_wf_detect' -s 450 -c 6 \
    -f -o gpurun_out/r02_final2_steady python tools/profile_step.py 2000000 1 > gpurun_out/r02_run16_ncu_steady.log 2>&1
tail -1 gpurun_out/r02_run16_ncu_steady.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sample_events|k_wf_root|k_wf_tree|k_wf_inject' -s 0 -c 4 \
    -f -o gpurun_out/r02_final2_side python tools/profile_step.py 2000000 1 > gpurun_out/r02_run16_ncu_side.log 2>&1
tail -1 gpurun_out/r02_run16_ncu_side.log
gzip -f gpurun_out/r02_final2_launches.csv
