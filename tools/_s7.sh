#!/bin/bash
mkdir -p gpurun_out
CONFIGS="instance refl" tools/ab_variants.sh base > gpurun_out/s7_ab.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "image_vs_reference or ray_counts or synthetic or batch_size or many_samples" > gpurun_out/s7_pytest.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/s7_refl_launches.csv python tools/profile_frame.py --frames 2 --config refl > gpurun_out/s7_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade --launch-skip 3 -c 2 -f -o gpurun_out/s7_refl_shade python tools/profile_frame.py --frames 2 --config refl > gpurun_out/s7_ncu2.log 2>&1
