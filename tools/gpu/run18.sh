#!/bin/bash
mkdir -p gpurun_out
export ART_SIDE_STREAM=0 ART_POOL_SLOTS=524288
timeout 300 python tools/profile_step.py 1000000 1 > gpurun_out/r02_run18_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sample_events' -s 1 -c 1 \
    -f -o gpurun_out/r02_run18_sampler python tools/profile_step.py 1000000 1 > gpurun_out/r02_run18_ncu.log 2>&1
tail -2 gpurun_out/r02_run18_ncu.log
