#!/bin/bash
# 1-GPU final set: GPU tests, configs 3 and 5 at size, the bench line (with the CPU baseline leg), ncu launch list + --set full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_run22_pytest.log
tail -3 gpurun_out/r02_run22_pytest.log
timeout 600 python tools/run_config.py --config 3 --events 10000000 --mass 1e-5 --tag c3 --out gpurun_out/r02_config3_1e7_1gpu_v3.json > gpurun_out/r02_run22_c3.log 2>&1; tail -c 300 gpurun_out/r02_run22_c3.log
timeout 900 python tools/run_config.py --config 5 --events 100000000 --no-fetch --tag c5s --out gpurun_out/r02_config5_1e8_1gpu_v3.json > gpurun_out/r02_run22_c5.log 2>&1; tail -c 300 gpurun_out/r02_run22_c5.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run22_bench.json 2> gpurun_out/r02_run22_bench.err; head -c 200 gpurun_out/r02_run22_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_run22_bench_ref.json 2> gpurun_out/r02_run22_bench_ref.err; head -c 300 gpurun_out/r02_run22_bench_ref.json
rm -rf gpurun_out/run_config
export ART_SIDE_STREAM=0 ART_POOL_SLOTS=524288
timeout 300 python tools/profile_step.py 2000000 1 > gpurun_out/r02_run22_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_final4_launches.csv python tools/profile_step.py 2000000 1 > gpurun_out/r02_run22_ncu_list.log 2>&1
tail -1 gpurun_out/r02_run22_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_wf_stages|k_wf_scan|k_wf_detect' -s 450 -c 6 \
    -f -o gpurun_out/r02_final4_steady python tools/profile_step.py 2000000 1 > gpurun_out/r02_run22_ncu_steady.log 2>&1
tail -1 gpurun_out/r02_run22_ncu_steady.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sample_events|k_wf_root|k_wf_tree|k_wf_inject' -s 0 -c 4 \
    -f -o gpurun_out/r02_final4_side python tools/profile_step.py 2000000 1 > gpurun_out/r02_run22_ncu_side.log 2>&1
tail -1 gpurun_out/r02_run22_ncu_side.log
gzip -f gpurun_out/r02_final4_launches.csv
