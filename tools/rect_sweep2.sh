#!/bin/bash
# development aid (run under gpurun): store flavours of the rect hot loop
for v in "" "-DTCW_RECT_SKEW(r_)=0" "-DTCW_RECT_ST(p_,v_)=__stcs(p_,v_)" "-DTCW_RECT_ST(p_,v_)=__stwt(p_,v_)"; do
  TCW_NVCC_EXTRA="$v" python -m pyfstat_b200.build --force > /dev/null 2>&1
  grep -c " error" pyfstat_b200/build.log
  for fl in fmn btsg; do
    echo -n "[$v]: "; python tools/prof_one.py rect 2880 64 "$fl" 5 2>&1 | tail -1 | sed "s/'table.*'map'/'map'/; s/'finalize.*//"
  done
done
python -m pyfstat_b200.build --force > /dev/null 2>&1
