#!/bin/bash
# timing experiments for k_screen_tc (production configuration: turns only): which of copies / MMAs / epilogue
# bounds the tile time.  debug bits: 2 = no bulk copies, 4 = no MMAs, 8 = no epilogue math, 32 = no Horner nodes
for dbg in ${DBGS:-0 6 8}; do
  echo "dbg=$dbg"
  NOPHI=1 PYATM_TC_SWAP=$dbg THETA=${THETA:-10} NS=8 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_screen_tc python tools/prof_tc.py 2>/dev/null | grep k_screen_tc | awk -F'","' '{print $NF}' | tail -1
done
