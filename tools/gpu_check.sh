#!/bin/bash
# full GPU test suite + in-graph kernel trace
python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -12
python tools/graph_trace.py 2>&1 | grep -v -i warn | head -${1:-24}
