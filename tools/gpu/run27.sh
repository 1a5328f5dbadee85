#!/bin/bash
# shipped build, 1 GPU: GPU tests, the bench line, ncu launch list + --set full captures (the files profiles/r02_ncu_summary.* and
# profiles/r02_bench_final.json are made from)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_run27_pytest.log
tail -2 gpurun_out/r02_run27_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run27_bench.json 2> gpurun_out/r02_run27_bench.err; head -c 200 gpurun_out/r02_run27_bench.json
export ART_SIDE_STREAM=0 ART_POOL_SLOTS=524288
timeout 300 python tools/profile_step.py 2000000 1 > gpurun_out/r02_run27_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r02_final5_launches.csv python tools/profile_step.py 2000000 1 > gpurun_out/r02_run27_ncu_list.log 2>&1
tail -1 gpurun_out/r02_run27_ncu_list.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_wf_stages|k_wf_scan|k_wf_detect' -s 450 -c 6 \
    -f -o gpurun_out/r02_final5_steady python tools/profile_step.py 2000000 1 > gpurun_out/r02_run27_ncu_steady.log 2>&1
tail -1 gpurun_out/r02_run27_ncu_steady.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample_events|k_wf_root|k_wf_tree|k_wf_inject' -s 0 -c 4 \
    -f -o gpurun_out/r02_final5_side python tools/profile_step.py 2000000 1 > gpurun_out/r02_run27_ncu_side.log 2>&1
tail -1 gpurun_out/r02_run27_ncu_side.log
gzip -f gpurun_out/r02_final5_launches.csv
