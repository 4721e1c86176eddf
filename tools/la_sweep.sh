#!/bin/bash
# potrf phase time for several look-ahead block widths (GOGP_LA_NB; 0 = one-stream recursion)
for n in ${LA_N:-32768}; do
for nb in ${LA_LIST:-0 2048 2048,0,128}; do
  echo "== N=$n GOGP_LA_NB=$nb"
  GOGP_LA_NB=$nb python tools/eval_once.py $n 2 2>&1 | tail -2
done
done
