#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_run11_pytest.log
tail -4 gpurun_out/r02_run11_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run11_bench.json 2> gpurun_out/r02_run11_bench.err; head -c 300 gpurun_out/r02_run11_bench.json
