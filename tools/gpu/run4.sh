#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r02_run4_pytest.log
tail -3 gpurun_out/r02_run4_pytest.log
ART_WF_TRACE=1 timeout 300 python tools/profile_step.py 4000000 1 > gpurun_out/r02_run4_trace4m.txt 2>&1; tail -1 gpurun_out/r02_run4_trace4m.txt
timeout 1500 python tools/variant_sweep.py 3000000 2 > gpurun_out/r02_run4_sweep.txt 2>&1
cat gpurun_out/r02_run4_sweep.txt
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run4_bench.json 2> gpurun_out/r02_run4_bench.err
head -c 250 gpurun_out/r02_run4_bench.json; tail -3 gpurun_out/r02_run4_bench.err
