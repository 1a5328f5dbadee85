#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/diff_variants.py default default@ART_POOL_SLOTS=262144,ART_SAMPLE_BLOCK=100000 2000000 > gpurun_out/r02_run14_diff_a.txt 2>&1
head -8 gpurun_out/r02_run14_diff_a.txt | cut -c1-400
timeout 900 python tools/diff_variants.py default adiabatic_raytracer_b200/variants/libart_old.so 2000000 > gpurun_out/r02_run14_diff_b.txt 2>&1
head -8 gpurun_out/r02_run14_diff_b.txt | cut -c1-400
