#!/bin/bash
# compute-sanitizer passes over the GPU test-suite (run under gpurun); summary -> gpurun_out/sanitizer_*.log
S=/usr/local/cuda/bin/compute-sanitizer
$S --tool memcheck python -m pytest tests -m gpu -x -q -k 'not full_size and not mcmc_step' > gpurun_out/sanitizer_memcheck_r2.log 2>&1
$S --tool memcheck python -m pytest tests -m gpu -x -q -k 'full_size_properties or mcmc_step' > gpurun_out/sanitizer_memcheck_full_r2.log 2>&1
$S --tool racecheck python -m pytest tests -m gpu -x -q -k 'tiled_kernels or golden or tie_rule or locate_pass or grid_search' > gpurun_out/sanitizer_racecheck_r2.log 2>&1
for f in gpurun_out/sanitizer_memcheck_r2.log gpurun_out/sanitizer_memcheck_full_r2.log gpurun_out/sanitizer_racecheck_r2.log; do echo "--- $f"; grep -v '^$' $f | tail -n 3; done
