#!/bin/bash
# round 2, GPU call 45: GPU suite + smoke() at the final HEAD
O=gpurun_out/r02_au
mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
