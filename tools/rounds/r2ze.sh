#!/bin/bash
# round 2, call ze: full GPU suite + smoke of the current build
mkdir -p gpurun_out
T=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > $T/r2ze_pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -6 $T/r2ze_pytest_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $T/r2ze_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $T/r2ze_smoke.log | cut -c1-300
