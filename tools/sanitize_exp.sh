#!/bin/bash
# compute-sanitizer over the exponential window's recurrence / tensor-core path (run under gpurun)
S=/usr/local/cuda/bin/compute-sanitizer
K='(exp_recurrence or tiled_kernels_within or exact_exp_mode or tie_rule or golden or tensor_pass or several_atoms or smaller_than) and not full_size'
timeout 1500 $S --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/sanitizer_memcheck_exp_r02_v4.log 2>&1
timeout 1500 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exp_recurrence_ragged or tiled_kernels_within or several_atoms" > gpurun_out/sanitizer_racecheck_exp_r02_v4.log 2>&1
for f in gpurun_out/sanitizer_memcheck_exp_r02_v4.log gpurun_out/sanitizer_racecheck_exp_r02_v4.log; do echo "--- $f"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $f | tail -n 4; done
