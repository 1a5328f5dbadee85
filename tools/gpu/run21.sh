#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_run21_pytest.log
tail -3 gpurun_out/r02_run21_pytest.log
python tools/ab_variants.py new=default prev=adiabatic_raytracer_b200/variants/libart_prev.so
python tools/diff_variants.py default adiabatic_raytracer_b200/variants/libart_prev.so 1000000 2>&1 | head -6 | cut -c1-300
