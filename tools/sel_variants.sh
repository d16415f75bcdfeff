#!/bin/bash
# phase timings of the selection kernels for several builds (JX_TRACE synchronises around every phase)
HITS=${HITS:-20000000}
for lib in "$@"; do
  echo "== $lib"
  JCVI_B200_LIB=$PWD/$lib JX_TRACE=1 python tools/step_probe.py $HITS 2>&1 | sed -n "/---- last step ----/,\$p" | grep "k_select\|select_cell\|per step"
done
