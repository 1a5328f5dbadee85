#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_run6_pytest.log
tail -4 gpurun_out/r02_run6_pytest.log
timeout 600 python tools/run_config.py --config 3 --events 10000000 --mass 1e-5 --tag c3 --out gpurun_out/r02_config3_1e7_1gpu.json > gpurun_out/r02_run6_c3.log 2>&1; tail -c 600 gpurun_out/r02_run6_c3.log
timeout 900 python tools/run_config.py --config 5 --events 100000000 --no-fetch --tag c5s --out gpurun_out/r02_config5_1e8_1gpu.json > gpurun_out/r02_run6_c5.log 2>&1; tail -c 600 gpurun_out/r02_run6_c5.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run6_bench.json 2> gpurun_out/r02_run6_bench.err; head -c 200 gpurun_out/r02_run6_bench.json
rm -rf gpurun_out/run_config
