#!/bin/bash
# round-2 final profiles of the exponential window's recurrence + tensor-core path (run under gpurun)
V=${V:-v4}
set -x
ncu --set full --clock-control none --import-source on -k regex:tcw_exptc_map -s 1 -c 1 -o gpurun_out/r02_exptc_${V} python tools/exp_prof.py 5760 8 > gpurun_out/prof_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcw_exp_walk -s 1 -c 1 -o gpurun_out/r02_expwalk_lut_${V} python tools/exp_prof.py 5760 8 > gpurun_out/prof_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcw_exp_walk -s 1 -c 1 -o gpurun_out/r02_expwalk_exact_${V} python tools/exp_prof.py 5760 8 exact > gpurun_out/prof_c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_${V}.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_${V}.json 2> gpurun_out/r02_bench_${V}.err
tail -2 gpurun_out/r02_bench_${V}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_${V}_reference.json 2>> gpurun_out/r02_bench_${V}.err
python tools/accuracy_report.py > gpurun_out/r02_accuracy_report_${V}.jsonl 2>> gpurun_out/r02_bench_${V}.err
ls -la gpurun_out/r02_*
