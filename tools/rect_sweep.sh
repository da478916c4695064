#!/bin/bash
# development aid (run under gpurun): time the rect map kernel (planner defaults, optional overrides via env)
for cfg in "2880 64" "2880 8" "1440 64" "5760 16"; do set -- $cfg
 for fl in "" fmn btsg; do
   python tools/prof_one.py rect $1 $2 "$fl" 5 2>&1 | tail -1 | sed "s/'table.*'map'/'map'/; s/'finalize.*//"
 done
done
