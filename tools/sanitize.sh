#!/bin/bash
# compute-sanitizer passes over the GPU test-suite (run under gpurun); summary -> gpurun_out/sanitizer_*.log
S=/usr/local/cuda/bin/compute-sanitizer
K='not full_size and not mcmc_step and not dist_gpu and not random_window_sweep'
$S --tool memcheck python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/sanitizer_memcheck_r02.log 2>&1
$S --tool racecheck python -m pytest tests -m gpu -x -q -k 'tiled_kernels or golden or locate_pass or grid_search or persistent_kernel or exp_tiled_row_classes or exp_lut_geometry_is_runtime' > gpurun_out/sanitizer_racecheck_r02.log 2>&1
$S --tool synccheck python -m pytest tests -m gpu -x -q -k 'persistent_kernel or tiled_kernels_within' > gpurun_out/sanitizer_synccheck_r02.log 2>&1
for f in gpurun_out/sanitizer_memcheck_r02.log gpurun_out/sanitizer_racecheck_r02.log gpurun_out/sanitizer_synccheck_r02.log; do echo "--- $f"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $f | tail -n 4; done
