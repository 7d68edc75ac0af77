#!/bin/bash
# timing experiments for k_screen_tc: which of copies / MMAs / epilogue bounds the tile time
for dbg in ${DBGS:-0 6 8}; do
  echo "dbg=$dbg"
  PYATM_TC_SWAP=$dbg THETA=${THETA:-6} NS=8 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_screen_tc python tools/prof_tc.py 2>/dev/null | grep k_screen_tc | awk -F'","' '{print $NF}' | tail -2
done
