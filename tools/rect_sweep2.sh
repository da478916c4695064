#!/bin/bash
# development aid (run under gpurun): rows-per-thread / occupancy variants of the rect kernel
for v in "2 1608 1472 4" "3 1100 960 2" "4 768 608 2" "2 1608 1472 2"; do set -- $v
  TCW_NVCC_EXTRA="-DTCW_RECT_MINB=$1 -DTCW_RECT_ECAP=$2 -DTCW_RECT_DT=$3" python -m pyfstat_b200.build --force > /dev/null 2>&1
  for fl in "" fmn btsg; do
    echo -n "MINB=$1 ECAP=$2 DT=$3 R=$4: "; TCW_RECT_R=$4 python tools/prof_one.py rect 2880 64 "$fl" 5 2>&1 | tail -1 | sed "s/'table.*'map'/'map'/; s/'finalize.*//"
  done
done
python -m pyfstat_b200.build --force > /dev/null 2>&1
