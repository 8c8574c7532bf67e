#!/bin/bash
# round 2: last check of the committed state -- smoke() and the full GPU test suite
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r02_tests_final2.log 2>&1
cat gpurun_out/r02_tests_final2.log
