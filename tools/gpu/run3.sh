#!/bin/bash
# GPU run 3 of round 2: tests, bench, ncu launch list (time + DRAM bytes per launch) and --set full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_run3_pytest.log
tail -3 gpurun_out/r02_run3_pytest.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run3_bench.json 2> gpurun_out/r02_run3_bench.err
head -c 300 gpurun_out/r02_run3_bench.json; tail -3 gpurun_out/r02_run3_bench.err
export ART_SIDE_STREAM=0 ART_POOL_SLOTS=524288
timeout 300 python tools/profile_step.py 2000000 1 > gpurun_out/r02_run3_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches.csv python tools/profile_step.py 2000000 1 > gpurun_out/r02_run3_ncu_list.log 2>&1
tail -2 gpurun_out/r02_run3_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_wf_stages|k_wf_scan|k_wf_detect' -s 450 -c 6 \
    -f -o gpurun_out/r02_steady python tools/profile_step.py 2000000 1 > gpurun_out/r02_run3_ncu_steady.log 2>&1
tail -2 gpurun_out/r02_run3_ncu_steady.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sample_events|k_wf_root|k_wf_tree|k_wf_inject' -s 0 -c 4 \
    -f -o gpurun_out/r02_side python tools/profile_step.py 2000000 1 > gpurun_out/r02_run3_ncu_side.log 2>&1
tail -2 gpurun_out/r02_run3_ncu_side.log
gzip -f gpurun_out/r02_launches.csv
ls -la gpurun_out | tail -12
