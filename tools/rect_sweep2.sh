#!/bin/bash
# development aid (run under gpurun): staged-slice capacity variants of the rect kernel
for v in "1608 1472" "1100 960" "860 704" "2200 2048"; do set -- $v
  TCW_NVCC_EXTRA="-DTCW_RECT_ECAP=$1 -DTCW_RECT_DT=$2" python -m pyfstat_b200.build --force > /dev/null 2>&1
  for fl in "" fmn btsg; do
    echo -n "ECAP=$1 DT=$2: "; python tools/prof_one.py rect 2880 64 "$fl" 5 2>&1 | tail -1 | sed "s/'table.*'map'/'map'/; s/'finalize.*//"
  done
done
python -m pyfstat_b200.build --force > /dev/null 2>&1
