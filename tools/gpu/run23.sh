#!/bin/bash
# 8-GPU box, shipped build: multi-GPU correctness at world 8, BASELINE config 5 (1e8 events, one job), 8-rank bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tools/check_multi_gpu.py 400000 > gpurun_out/r02_check_multi_gpu_8_v3.json 2> gpurun_out/r02_run23_check.err; cat gpurun_out/r02_check_multi_gpu_8_v3.json | cut -c1-300
timeout 600 $TR --master-port 29522 tools/run_config.py --config 5 --events 100000000 --tag c5 --out gpurun_out/r02_config5_1e8_8gpu_v3.json > gpurun_out/r02_run23_c5.log 2>&1; tail -c 300 gpurun_out/r02_run23_c5.log
timeout 400 $TR --master-port 29525 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_run23_bench8.json 2> gpurun_out/r02_run23_bench8.err; head -c 220 gpurun_out/r02_run23_bench8.json; tail -2 gpurun_out/r02_run23_bench8.err
rm -rf gpurun_out/run_config gpurun_out/check_multi
