#!/bin/bash
# 2-GPU box: e2e breakdown on one GPU, 2-rank bench with hist_check, multi-GPU product entry point against the single-GPU file
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_run5_bench1.json 2> gpurun_out/r02_run5_bench1.err
python -c "import json; d=json.load(open('gpurun_out/r02_run5_bench1.json')); print('1 GPU value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), d['e2e']['parts_rank0_s'], d['hist_check'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_run5_bench2.json 2> gpurun_out/r02_run5_bench2.err
python -c "import json; d=json.load(open('gpurun_out/r02_run5_bench2.json')); print('2 GPU value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), d['hist_check'])"; tail -3 gpurun_out/r02_run5_bench2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py 200000 > gpurun_out/r02_check_multi_gpu_2.json 2> gpurun_out/r02_run5_check.err
cat gpurun_out/r02_check_multi_gpu_2.json; tail -5 gpurun_out/r02_run5_check.err
rm -rf gpurun_out/check_multi
