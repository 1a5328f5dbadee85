#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/diff_variants.py default adiabatic_raytracer_b200/variants/libart_old.so 2000000 > gpurun_out/r02_run13_diff.txt 2>&1
tail -30 gpurun_out/r02_run13_diff.txt | cut -c1-600
