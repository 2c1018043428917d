#!/bin/bash
# A/B on ONE box: usage gpu_ab.sh ENVVAR  -> step span with ENVVAR=0 and unset, twice each (interleaved)
for i in 1 2; do
  echo "== $1=0"; env $1=0 python tools/graph_trace.py 2>&1 | grep -E "span|gelu|<256, false, true, false, 2" | head -5
  echo "== default"; python tools/graph_trace.py 2>&1 | grep -E "span|gelu|<256, false, true, false, 2" | head -5
done
